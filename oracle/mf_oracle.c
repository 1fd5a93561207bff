/* TEST INFRASTRUCTURE - CPU oracle (see mf_oracle.h for scope, provenance and parity status).
 *
 * Every function restates one piece of the reference, cited as file:line under
 * /root/reference/src.  The float/double promotions of the reference's C++ expressions are
 * written out explicitly, because the test that pins this file (tests/test_oracle_vs_ref.py)
 * demands BIT-equal results against the reference's own sources built with -ffp-contract=off.
 * Build: oracle/Makefile (gcc -std=c11 -O2 -ffp-contract=off).
 */
#define _GNU_SOURCE
#include "mf_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ---------------- BLAS-1 as oracle/shim/mkl.h defines it: sequential fp32, index order ------ */
static float o_sdot(int n, const float* x, const float* y) {
  float acc = 0.0f;
  for (int i = 0; i < n; i++) acc += x[i] * y[i];
  return acc;
}
static void o_saxpy(int n, float a, const float* x, float* y) {
  for (int i = 0; i < n; i++) y[i] += a * x[i];
}

/* util.h:163-165, CACHE_LINE_SIZE = 64 (util.h:37-41) */
int mfo_padding(int dim) {
  return (int)((((size_t)dim * sizeof(float) - 1) / 64 * 64 + 64) / sizeof(float));
}

/* model.cc:36-38:  eta_ = (float)(eta0_ * 1.0/pow(round,gam_)); */
float mfo_seteta(float eta0, int round, float gam) {
  return (float)((double)eta0 * 1.0 / pow((double)round, (double)gam));
}
/* model.cc:350-352 */
float mfo_seteta_cutoff(float eta0, int round, float gam, float mineta) {
  float e = mfo_seteta(eta0, round, gam);
  return mineta > e ? mineta : e; /* std::max(mineta_, e) returns mineta_ unless mineta_ < e */
}

void mfo_srand(unsigned seed) { srand(seed); }
/* n draws of the reference's validation index, `rand() % nvalid` (admf.h:82), in the order the filter makes them
 * (one per user of the file) */
void mfo_rand_draws(int64_t n, int64_t nvalid, int32_t* out) {
  for (int64_t i = 0; i < n; i++) out[i] = (int32_t)((size_t)rand() % (size_t)nvalid);
}


/* =============================== blocks.proto wire format =================================== */
static int get_varint(const uint8_t** pp, const uint8_t* end, uint64_t* out) {
  const uint8_t* p = *pp;
  uint64_t v = 0;
  for (int shift = 0; shift < 70 && p < end; shift += 7) {
    uint8_t b = *p++;
    if (shift < 64) v |= (uint64_t)(b & 0x7F) << shift;
    if (!(b & 0x80)) {
      *out = v;
      *pp = p;
      return 1;
    }
  }
  return 0;
}
static int skip_field(unsigned wt, const uint8_t** pp, const uint8_t* end) {
  uint64_t tmp;
  switch (wt) {
    case 0: return get_varint(pp, end, &tmp);
    case 1: if (end - *pp < 8) return 0; *pp += 8; return 1;
    case 2: if (!get_varint(pp, end, &tmp) || (uint64_t)(end - *pp) < tmp) return 0; *pp += tmp; return 1;
    case 5: if (end - *pp < 4) return 0; *pp += 4; return 1;
    default: return 0;
  }
}

typedef struct {
  int64_t *block_off, *run_off;
  int32_t *run_uid, *vid;
  float* rating;
  int64_t nblocks, nruns, nrat;
  int64_t cap_block_off, cap_run_off, cap_run_uid, cap_vid, cap_rating; /* one per array */
} builder;

static void* grow(void* p, int64_t* cap, int64_t need, size_t elt) {
  if (need <= *cap) return p;
  int64_t nc = *cap ? *cap : 1024;
  while (nc < need) nc *= 2;
  *cap = nc;
  return realloc(p, (size_t)nc * elt);
}

static int parse_record(builder* b, const uint8_t* p, const uint8_t* end) {
  int32_t vid = 0;
  float rating = 0.f;
  while (p < end) {
    uint64_t tag, v;
    if (!get_varint(&p, end, &tag)) return 0;
    if (tag == 0x08) { /* required int32 vid = 1 */
      if (!get_varint(&p, end, &v)) return 0;
      vid = (int32_t)(uint32_t)v;
    } else if (tag == 0x15) { /* required float rating = 2 */
      if (end - p < 4) return 0;
      memcpy(&rating, p, 4);
      p += 4;
    } else if (!skip_field((unsigned)(tag & 7), &p, end)) {
      return 0;
    }
  }
  b->vid = grow(b->vid, &b->cap_vid, b->nrat + 1, sizeof(int32_t));
  b->rating = grow(b->rating, &b->cap_rating, b->nrat + 1, sizeof(float));
  b->vid[b->nrat] = vid;
  b->rating[b->nrat] = rating;
  b->nrat++;
  return 1;
}

static int parse_user(builder* b, const uint8_t* p, const uint8_t* end) {
  b->run_uid = grow(b->run_uid, &b->cap_run_uid, b->nruns + 1, sizeof(int32_t));
  b->run_off = grow(b->run_off, &b->cap_run_off, b->nruns + 2, sizeof(int64_t));
  int64_t run = b->nruns++;
  b->run_uid[run] = 0;
  b->run_off[run] = b->nrat;
  while (p < end) {
    uint64_t tag, v;
    if (!get_varint(&p, end, &tag)) return 0;
    if (tag == 0x08) { /* required int32 uid = 1 */
      if (!get_varint(&p, end, &v)) return 0;
      b->run_uid[run] = (int32_t)(uint32_t)v;
    } else if (tag == 0x12) { /* repeated Record record = 2 */
      if (!get_varint(&p, end, &v) || (uint64_t)(end - p) < v) return 0;
      if (!parse_record(b, p, p + v)) return 0;
      p += v;
    } else if (!skip_field((unsigned)(tag & 7), &p, end)) {
      return 0;
    }
  }
  b->run_off[run + 1] = b->nrat;
  return 1;
}

static int parse_block(builder* b, const uint8_t* p, const uint8_t* end) {
  b->block_off = grow(b->block_off, &b->cap_block_off, b->nblocks + 2, sizeof(int64_t));
  b->block_off[b->nblocks] = b->nruns;
  while (p < end) {
    uint64_t tag, v;
    if (!get_varint(&p, end, &tag)) return 0;
    if (tag == 0x0A) { /* repeated User user = 1 */
      if (!get_varint(&p, end, &v) || (uint64_t)(end - p) < v) return 0;
      if (!parse_user(b, p, p + v)) return 0;
      p += v;
    } else if (!skip_field((unsigned)(tag & 7), &p, end)) {
      return 0;
    }
  }
  b->nblocks++;
  b->block_off[b->nblocks] = b->nruns;
  return 1;
}

/* util.h:76-88 (plain_read): while (fread(&isize,...)) { fread(buf, isize); ParseFromArray } */
mfo_file* mfo_read_blocks(const char* path) {
  FILE* f = fopen(path, "rb");
  if (!f) return NULL;
  builder b;
  memset(&b, 0, sizeof b);
  b.block_off = grow(b.block_off, &b.cap_block_off, 2, sizeof(int64_t));
  b.run_off = grow(b.run_off, &b.cap_run_off, 2, sizeof(int64_t));
  b.block_off[0] = 0;
  b.run_off[0] = 0;
  uint8_t* buf = NULL;
  size_t cap = 0;
  uint32_t isize;
  int ok = 1;
  while (fread(&isize, 1, sizeof isize, f) == sizeof isize) {
    if (isize > cap) {
      cap = isize;
      buf = realloc(buf, cap);
    }
    if (fread(buf, 1, isize, f) != isize || !parse_block(&b, buf, buf + isize)) {
      ok = 0;
      break;
    }
  }
  free(buf);
  fclose(f);
  if (!ok) {
    free(b.block_off); free(b.run_off); free(b.run_uid); free(b.vid); free(b.rating);
    return NULL;
  }
  mfo_file* out = calloc(1, sizeof *out);
  out->d.nblocks = b.nblocks;
  out->d.block_off = b.block_off;
  out->d.nruns = b.nruns;
  out->d.run_uid = b.run_uid;
  out->d.run_off = b.run_off;
  out->d.vid = b.vid;
  out->d.rating = b.rating;
  out->nratings = b.nrat;
  return out;
}

void mfo_free_file(mfo_file* f) {
  if (!f) return;
  free((void*)f->d.block_off); free((void*)f->d.run_off); free((void*)f->d.run_uid);
  free((void*)f->d.vid); free((void*)f->d.rating);
  free(f);
}

static int varint_size(uint64_t v) {
  int n = 1;
  while (v >= 0x80) { v >>= 7; n++; }
  return n;
}
static uint8_t* put_varint(uint8_t* p, uint64_t v) {
  while (v >= 0x80) { *p++ = (uint8_t)((v & 0x7F) | 0x80); v >>= 7; }
  *p++ = (uint8_t)v;
  return p;
}

/* getdata.cc:100-103: fwrite(&size,4); fwrite(bytes) per Block */
int mfo_write_blocks(const char* path, const mfo_data* d) {
  FILE* f = fopen(path, "wb");
  if (!f) return -1;
  uint8_t* buf = NULL;
  size_t cap = 0;
  for (int64_t b = 0; b < d->nblocks; b++) {
    size_t total = 0;
    for (int64_t r = d->block_off[b]; r < d->block_off[b + 1]; r++) {
      size_t us = 1 + varint_size((uint64_t)(int64_t)d->run_uid[r]);
      for (int64_t k = d->run_off[r]; k < d->run_off[r + 1]; k++)
        us += 2 + (1 + varint_size((uint64_t)(int64_t)d->vid[k]) + 5);
      total += 1 + varint_size(us) + us;
    }
    if (total > cap) { cap = total * 2; buf = realloc(buf, cap); }
    uint8_t* p = buf;
    for (int64_t r = d->block_off[b]; r < d->block_off[b + 1]; r++) {
      size_t us = 1 + varint_size((uint64_t)(int64_t)d->run_uid[r]);
      for (int64_t k = d->run_off[r]; k < d->run_off[r + 1]; k++)
        us += 2 + (1 + varint_size((uint64_t)(int64_t)d->vid[k]) + 5);
      *p++ = 0x0A;
      p = put_varint(p, us);
      *p++ = 0x08;
      p = put_varint(p, (uint64_t)(int64_t)d->run_uid[r]);
      for (int64_t k = d->run_off[r]; k < d->run_off[r + 1]; k++) {
        int rs = 1 + varint_size((uint64_t)(int64_t)d->vid[k]) + 5;
        *p++ = 0x12;
        *p++ = (uint8_t)rs; /* <= 16 */
        *p++ = 0x08;
        p = put_varint(p, (uint64_t)(int64_t)d->vid[k]);
        *p++ = 0x15;
        memcpy(p, &d->rating[k], 4);
        p += 4;
      }
    }
    uint32_t sz = (uint32_t)(p - buf);
    if (fwrite(&sz, 1, 4, f) != 4 || fwrite(buf, 1, sz, f) != sz) { fclose(f); free(buf); return -1; }
  }
  free(buf);
  return fclose(f);
}

/* ==================================== plain SGD ============================================ */
/* SgdFilter::operator(), mf.h:76-132.  The two loop bodies (mf.h:94-109 and 113-128) are
 * identical and the prefetch (mf.h:89-93) is compiled out (no -DFETCH in Makefile:23), so one
 * loop over the whole run restates both. */
void mfo_sgd_epoch(mfo_model* m, const mfo_data* d, float eta, float lambda, float gb) {
  const int dim = m->dim;
  float* q = malloc(sizeof(float) * (size_t)dim);
  for (int64_t b = 0; b < d->nblocks; b++) {
    /* mf.h:80  const float lameta = 1.0-mf_.eta_*mf_.lambda_;  (fp32 product, double subtract) */
    const float lameta = (float)(1.0 - (double)(eta * lambda));
    /* mf.h:104 passes the double (lameta-1.0) to a `const float alpha` parameter */
    const float lm1 = (float)((double)lameta - 1.0);
    for (int64_t r = d->block_off[b]; r < d->block_off[b + 1]; r++) {
      const int uid = d->run_uid[r];
      float* theta = m->theta + (size_t)uid * m->stride;
      for (int64_t k = d->run_off[r]; k < d->run_off[r + 1]; k++) {
        memset(q, 0, sizeof(float) * (size_t)dim);                    /* mf.h:94  */
        const int vid = d->vid[k];
        float* phi = m->phi + (size_t)vid * m->stride;
        const float rating = d->rating[k];
        float error = rating - o_sdot(dim, theta, phi) - m->bu[uid] - m->bv[vid] - gb; /* :99-101 */
        error = eta * error;                                          /* mf.h:102 */
        o_saxpy(dim, error, theta, q);                                /* mf.h:103 */
        o_saxpy(dim, lm1, theta, theta);                              /* mf.h:104 */
        o_saxpy(dim, error, phi, theta);                              /* mf.h:105 */
        o_saxpy(dim, lameta, phi, q);                                 /* mf.h:106 */
        memcpy(phi, q, sizeof(float) * (size_t)dim);                  /* mf.h:107 */
        m->bu[uid] = lameta * m->bu[uid] + error;                     /* mf.h:108 */
        m->bv[vid] = lameta * m->bv[vid] + error;                     /* mf.h:109 */
      }
    }
  }
  free(q);
}

/* ==================================== evaluation =========================================== */
/* MF::calc_mse, model.cc:41-73 */
float mfo_sse(const mfo_model* m, const mfo_data* d, float gb, int64_t* ndata) {
  float sloss = 0.0f;
  int64_t n = 0;
  for (int64_t b = 0; b < d->nblocks; b++) {
    float sl = 0.0f;
    for (int64_t r = d->block_off[b]; r < d->block_off[b + 1]; r++) {
      const int uid = d->run_uid[r];
      const float* theta = m->theta + (size_t)uid * m->stride;
      n += d->run_off[r + 1] - d->run_off[r];
      for (int64_t k = d->run_off[r]; k < d->run_off[r + 1]; k++) {
        const int vid = d->vid[k];
        const float* phi = m->phi + (size_t)vid * m->stride;
        float error = d->rating[k] - o_sdot(m->dim, theta, phi) - m->bu[uid] - m->bv[vid] - gb;
        sl += error * error; /* model.cc:64 */
      }
    }
    sloss += sl; /* model.cc:68 */
  }
  if (ndata) *ndata = n;
  return sloss;
}

static float active(float val, int type);

/* The evaluation that matches --loss: MF::calc_mse (model.cc:41-73) with the prediction sent through the link
 * active() of util.h:90-95 first, exactly as the admf update (admf.h:69) and updateReg (model.h:87-88) form
 * their residual: error = rating - active(sdot + bu + bv + gb, loss).  The reference's calc_mse itself ignores
 * loss_ (model.cc:62) and measure_ is never read (model.h:109), so this is the SURVEY 8f-4 extension; loss = 0
 * differs from mfo_sse only in the order of the additions. */
float mfo_sse_link(const mfo_model* m, const mfo_data* d, float gb, int loss, int64_t* ndata) {
  float sloss = 0.0f;
  int64_t n = 0;
  for (int64_t b = 0; b < d->nblocks; b++) {
    float sl = 0.0f;
    for (int64_t r = d->block_off[b]; r < d->block_off[b + 1]; r++) {
      const int uid = d->run_uid[r];
      const float* theta = m->theta + (size_t)uid * m->stride;
      n += d->run_off[r + 1] - d->run_off[r];
      for (int64_t k = d->run_off[r]; k < d->run_off[r + 1]; k++) {
        const int vid = d->vid[k];
        const float* phi = m->phi + (size_t)vid * m->stride;
        float error = d->rating[k] - active(o_sdot(m->dim, theta, phi) + m->bu[uid] + m->bv[vid] + gb, loss);
        sl += error * error;
      }
    }
    sloss += sl;
  }
  if (ndata) *ndata = n;
  return sloss;
}

/* ==================================== SGLD / DP ============================================ */
/* model.cc:240-242 */
float mfo_dp_bound(float epsilon, int tau) {
  if (epsilon <= 0.0f) return 1.0f;
  return (float)((double)epsilon * 1.0 / (4.0 * 25.0 * (double)tau));
}

/* model.cc:247-297: block_count + ur_/vr_ */
int32_t mfo_dp_weights(const mfo_data* d, int nu, int nv, float* ur, float* vr) {
  int* uc = calloc((size_t)nu, sizeof(int));
  int* vc = calloc((size_t)nv, sizeof(int));
  int32_t ntrain = 0;
  for (int64_t r = 0; r < d->nruns; r++)
    for (int64_t k = d->run_off[r]; k < d->run_off[r + 1]; k++) {
      uc[d->run_uid[r]] += 1;
      vc[d->vid[k]] += 1;
      ntrain++;
    }
  for (int i = 0; i < nu; i++) ur[i] = (float)ntrain / uc[i]; /* model.cc:294 (inf if unseen) */
  for (int i = 0; i < nv; i++) vr[i] = (float)ntrain / vc[i];
  free(uc);
  free(vc);
  return ntrain;
}

void mfo_noise_from_table(void* ctx, int kind, int32_t row, int64_t t, int32_t j, int32_t dim,
                          float* out) {
  (void)kind; (void)row; (void)t;
  const mfo_noise_table* nt = ctx;
  /* dpmf.h:53-54 draw thetaind/phiind per user, dpmf.h:87 advances both by dim+1 per record;
   * model.cc:316,324 draw a fresh rndind per row (j == -1 here). */
  int64_t ind = (int64_t)nt->offset + (j < 0 ? 0 : (int64_t)j * (dim + 1));
  memcpy(out, nt->table + ind, sizeof(float) * (size_t)(dim + 1));
}

/* Philox4x32-10 (Salmon et al., SC'11; the constants are those of Random123 / cuRAND). */
void mfo_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
  uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
  uint32_t k0 = key[0], k1 = key[1];
  for (int round = 0; round < 10; round++) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    uint32_t n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* DESIGN.md "SGLD noise stream": counter = (t, row, chunk, kind + 2*round), key = seed.
 * Box-Muller on the top 24 bits of each word: u1 = ((x>>8)+1)/2^24 in (0,1], u2 = (x>>8)/2^24 in [0,1). */
static void philox_block(uint64_t seed, uint32_t round, int kind, int32_t row, int64_t t,
                         uint32_t chunk, uint32_t x[4]) {
  uint32_t ctr[4] = {(uint32_t)t, (uint32_t)row, chunk, (uint32_t)kind + 2u * round};
  uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
  mfo_philox4x32_10(ctr, key, x);
}

void mfo_philox_normal4(uint64_t seed, uint32_t round, int kind, int32_t row, int64_t t,
                        uint32_t chunk, float out[4]) {
  uint32_t x[4];
  philox_block(seed, round, kind, row, t, chunk, x);
  for (int p = 0; p < 2; p++) {
    float u1 = ((float)(x[2 * p] >> 8) + 1.0f) * (1.0f / 16777216.0f);
    float u2 = (float)(x[2 * p + 1] >> 8) * (1.0f / 16777216.0f);
    float rad = sqrtf(-2.0f * logf(u1));
    float ang = 6.28318530717958647692f * u2;
    out[2 * p] = rad * cosf(ang);
    out[2 * p + 1] = rad * sinf(ang);
  }
}

/* The bias value of (kind,row,t) is made from the bits the transforms above leave unused: the low
 * bytes of words 0..2 of chunk 0 are the 24-bit u1, those of chunk 1 the 24-bit u2; z = r cos. */
float mfo_philox_bias_normal(uint64_t seed, uint32_t round, int kind, int32_t row, int64_t t) {
  uint32_t x0[4], x1[4];
  philox_block(seed, round, kind, row, t, 0u, x0);
  philox_block(seed, round, kind, row, t, 1u, x1);
  uint32_t s0 = (x0[0] & 0xFFu) | ((x0[1] & 0xFFu) << 8) | ((x0[2] & 0xFFu) << 16);
  uint32_t s1 = (x1[0] & 0xFFu) | ((x1[1] & 0xFFu) << 8) | ((x1[2] & 0xFFu) << 16);
  float u1 = ((float)s0 + 1.0f) * (1.0f / 16777216.0f);
  float u2 = (float)s1 * (1.0f / 16777216.0f);
  return sqrtf(-2.0f * logf(u1)) * cosf(6.28318530717958647692f * u2);
}

void mfo_noise_from_philox(void* ctx, int kind, int32_t row, int64_t t, int32_t j, int32_t dim,
                           float* out) {
  (void)j;
  const mfo_noise_philox* ph = ctx;
  float z[4];
  for (int c = 0; c < dim; c += 4) {
    mfo_philox_normal4(ph->seed, ph->round, kind, row, t, (uint32_t)(c / 4), z);
    for (int i = 0; i < 4 && c + i < dim; i++) out[c + i] = z[i];
  }
  out[dim] = mfo_philox_bias_normal(ph->seed, ph->round, kind, row, t);
}

/* SgldFilter::operator(), dpmf.h:41-91 */
void mfo_sgld_epoch(mfo_model* m, const mfo_data* d, mfo_dp_state* st, float gb,
                    mfo_noise_fn noise, void* noise_ctx) {
  const int dim = m->dim;
  float* q = malloc(sizeof(float) * (size_t)dim);
  float* p = malloc(sizeof(float) * (size_t)dim);
  float* xu = malloc(sizeof(float) * (size_t)(dim + 1));
  float* xv = malloc(sizeof(float) * (size_t)(dim + 1));
  for (int64_t b = 0; b < d->nblocks; b++) {
    const float eta = st->eta;                                              /* dpmf.h:45 */
    const float scal = eta * st->ntrain * st->bound * st->lambda_r;         /* dpmf.h:46 */
    for (int64_t r = d->block_off[b]; r < d->block_off[b + 1]; r++) {
      const int uid = d->run_uid[r];
      float* theta = m->theta + (size_t)uid * m->stride;
      int32_t j = 0;
      for (int64_t k = d->run_off[r]; k < d->run_off[r + 1]; k++, j++) {
        memset(q, 0, sizeof(float) * (size_t)dim);
        const int vid = d->vid[k];
        float* phi = m->phi + (size_t)vid * m->stride;
        const float rating = d->rating[k];
        /* dpmf.h:61-66: int gc,vc,uc receive uint64 arithmetic truncated to int */
        int gc = (int)(st->gcount++);
        int vc = (int)((uint64_t)(int64_t)gc - st->gcountv[vid]);
        st->gcountv[vid] = (uint64_t)(int64_t)gc;
        int uc = (int)((uint64_t)(int64_t)gc - st->gcountu[uid]);
        st->gcountu[uid] = (uint64_t)(int64_t)gc;
        noise(noise_ctx, 0, uid, gc, j, dim, xu);
        noise(noise_ctx, 1, vid, gc, j, dim, xv);
        const float su = sqrtf(st->temp * eta * uc); /* sqrt(float) -> float overload */
        const float sv = sqrtf(st->temp * eta * vc);
        o_saxpy(dim, su, xu, theta);                                         /* dpmf.h:67 */
        o_saxpy(dim, sv, xv, phi);                                           /* dpmf.h:68 */
        m->bu[uid] += su * xu[dim];                                          /* dpmf.h:69 */
        m->bv[vid] += sv * xv[dim];                                          /* dpmf.h:70 */
        float error = rating - o_sdot(dim, theta, phi) - m->bu[uid] - m->bv[vid] - gb; /* :72-74 */
        error = scal * error;                                                /* dpmf.h:75 */
        o_saxpy(dim, error, theta, q);                                       /* dpmf.h:76 */
        for (int i = 0; i < dim; i++) p[i] = st->lambda_u[i] * theta[i];     /* dpmf.h:77 */
        o_saxpy(dim, -eta * st->ur[uid] * st->bound, p, theta);              /* dpmf.h:78 */
        o_saxpy(dim, error, phi, theta);                                     /* dpmf.h:79 */
        for (int i = 0; i < dim; i++) p[i] = st->lambda_v[i] * phi[i];       /* dpmf.h:80 */
        o_saxpy(dim, -eta * st->vr[vid] * st->bound, p, phi);                /* dpmf.h:81 */
        o_saxpy(dim, 1.0f, q, phi);                                          /* dpmf.h:82 */
        /* dpmf.h:84-85: (1.0 - fp32 product) evaluated in double, whole RHS in double */
        m->bu[uid] = (float)((1.0 - (double)(eta * st->lambda_ub * st->ur[uid] * st->bound)) *
                                 (double)m->bu[uid] + (double)error);
        m->bv[vid] = (float)((1.0 - (double)(eta * st->lambda_vb * st->vr[vid] * st->bound)) *
                                 (double)m->bv[vid] + (double)error);
      }
    }
  }
  free(q); free(p); free(xu); free(xv);
}

/* DPMF::finish_noise, model.cc:312-332 */
void mfo_finish_noise(mfo_model* m, mfo_dp_state* st, mfo_noise_fn noise, void* noise_ctx) {
  const int dim = m->dim;
  const int gc = (int)st->gcount;
  float* x = malloc(sizeof(float) * (size_t)(dim + 1));
  for (int i = 0; i < m->nu; i++) {
    int uc = (int)((uint64_t)(int64_t)gc - st->gcountu[i]);
    st->gcountu[i] = 0;
    noise(noise_ctx, 0, i, gc, -1, dim, x);
    const float sc = sqrtf(st->temp * st->eta * uc);
    o_saxpy(dim, sc, x, m->theta + (size_t)i * m->stride);
    m->bu[i] += sc * x[dim];
  }
  for (int i = 0; i < m->nv; i++) {
    int vc = (int)((uint64_t)(int64_t)gc - st->gcountv[i]);
    st->gcountv[i] = 0;
    noise(noise_ctx, 1, i, gc, -1, dim, x);
    const float sc = sqrtf(st->temp * st->eta * vc);
    o_saxpy(dim, sc, x, m->phi + (size_t)i * m->stride);
    m->bv[i] += sc * x[dim];
  }
  st->gcount = 0;
  free(x);
}

/* util.h:103-109 */
static float next_float(void) { return (float)((double)(float)rand() / ((double)(float)RAND_MAX + 1.0)); }
static float next_float2(void) {
  return (float)(((double)(float)rand() + 1.0) / ((double)(float)RAND_MAX + 2.0));
}
/* util.h:115-124 (polar Box-Muller).  2*next_float2() is fp32, "- 1.0" is double; log(s) and
 * sqrt pick the overload of their argument type: logf(s) (fp32), sqrt(double). */
static float sample_normal(void) {
  float x, y, s;
  do {
    x = (float)((double)(2 * next_float2()) - 1.0);
    y = (float)((double)(2 * next_float2()) - 1.0);
    s = x * x + y * y;
  } while (s >= 1.0 || s == 0.0);
  return (float)((double)x * sqrt(-2.0 * (double)logf(s) / (double)s));
}
/* util.h:126-148 (Marsaglia-Tsang) */
float mfo_sample_gamma(float alpha, float beta) {
  if (alpha < 1.0) {
    float u;
    do { u = next_float(); } while (u == 0.0);
    return (float)((double)mfo_sample_gamma((float)((double)alpha + 1.0), beta) *
                   pow((double)u, 1.0 / (double)alpha));
  } else {
    float d, c, x, v, u;
    d = (float)((double)alpha - 1.0 / 3.0);
    c = (float)(1.0 / sqrt(9.0 * (double)d));
    do {
      do {
        x = sample_normal();
        v = (float)(1.0 + (double)(c * x));
      } while (v <= 0.0);
      v = v * v * v;
      u = next_float();
    } while (((double)u >= (1.0 - 0.0331 * (double)(x * x) * (double)(x * x))) &&
             ((double)logf(u) >= (0.5 * (double)x * (double)x + (double)d * (1.0 - (double)v + (double)logf(v)))));
    return d * v / beta;
  }
}
/* util.h:150-154 */
static void gamma_posterior(float* lambda, float prior_alpha, float prior_beta, float psum_sqr,
                            float psum_cnt) {
  float alpha = (float)((double)prior_alpha + 0.5 * (double)psum_cnt);
  float beta = (float)((double)prior_beta + 0.5 * (double)psum_sqr);
  *lambda = mfo_sample_gamma(alpha, beta);
}

/* DPMF::sample_hyper, model.cc:335-348.  int arguments (ntrain_, nu_, nv_) convert to the float
 * parameter psum_cnt. */
void mfo_sample_hyper(const mfo_model* m, mfo_dp_state* st, float hyper_a, float hyper_b,
                      float train_sse) {
  const int dim = m->dim;
  gamma_posterior(&st->lambda_r, hyper_a, hyper_b, train_sse, (float)st->ntrain);
  gamma_posterior(&st->lambda_ub, hyper_a, hyper_b, o_sdot(m->nu, m->bu, m->bu), (float)m->nu);
  gamma_posterior(&st->lambda_vb, hyper_a, hyper_b, o_sdot(m->nv, m->bv, m->bv), (float)m->nv);
  float* normu = calloc((size_t)dim, sizeof(float));
  float* normv = calloc((size_t)dim, sizeof(float));
  /* util.h:156-161 normsqr_col: norm[i] += m[j][i]*m[j][i], j ascending */
  for (int i = 0; i < dim; i++) {
    for (int j = 0; j < m->nu; j++) {
      float x = m->theta[(size_t)j * m->stride + i];
      normu[i] += x * x;
    }
    for (int j = 0; j < m->nv; j++) {
      float x = m->phi[(size_t)j * m->stride + i];
      normv[i] += x * x;
    }
  }
  for (int i = 0; i < dim; i++) {
    gamma_posterior(&st->lambda_u[i], hyper_a, hyper_b, normu[i], (float)m->nu);
    gamma_posterior(&st->lambda_v[i], hyper_a, hyper_b, normv[i], (float)m->nv);
  }
  free(normu);
  free(normv);
}

/* ==================================== adaptive regulariser ================================= */
/* util.h:90-95 */
static float active(float val, int type) {
  if (type == 1) return 1.0f / (1.0f + expf(-val));
  return val;
}

/* model.cc:413 std::random_shuffle(begin,end) as libstdc++ (bits/stl_algo.h) implements it */
void mfo_shuffle_valid(int64_t n, int32_t* u, int32_t* v, float* r) {
  for (int64_t i = 1; i < n; i++) {
    int64_t j = rand() % (i + 1);
    if (i != j) {
      int32_t tu = u[i]; u[i] = u[j]; u[j] = tu;
      int32_t tv = v[i]; v[i] = v[j]; v[j] = tv;
      float tr = r[i]; r[i] = r[j]; r[j] = tr;
    }
  }
}

/* AdaptRegMF::updateReg / updateUV / updateBias, model.h:86-102 */
static void update_reg(mfo_model* m, mfo_ad_state* st, float gb, int uid, int vid, float rating) {
  const int dim = m->dim;
  const float* theta = m->theta + (size_t)uid * m->stride;
  const float* phi = m->phi + (size_t)vid * m->stride;
  float pred = active(o_sdot(dim, theta, phi) + m->bu[uid] + m->bv[vid] + gb, st->loss);
  float grad = rating - pred; /* util.h:96-101: both loss types */
  float inner = o_sdot(dim, st->theta_old + (size_t)uid * m->stride, phi);
  float t = st->lam_u - st->eta_reg * st->eta * grad * inner;
  st->lam_u = 0.0f < t ? t : 0.0f; /* std::max(0.0f, t) */
  inner = o_sdot(dim, theta, st->phi_old + (size_t)vid * m->stride);
  t = st->lam_v - st->eta_reg * st->eta * grad * inner;
  st->lam_v = 0.0f < t ? t : 0.0f;
  t = st->lam_bu - st->eta_reg * st->eta * grad * st->bu_old[uid];
  st->lam_bu = 0.0f < t ? t : 0.0f;
  t = st->lam_bv - st->eta_reg * st->eta * grad * st->bv_old[vid];
  st->lam_bv = 0.0f < t ? t : 0.0f;
}

/* AdRegFilter::operator(), admf.h:52-86 */
void mfo_admf_epoch(mfo_model* m, const mfo_data* d, mfo_ad_state* st, float gb) {
  const int dim = m->dim;
  float* q = malloc(sizeof(float) * (size_t)dim);
  for (int64_t b = 0; b < d->nblocks; b++) {
    const float eta = st->eta; /* admf.h:56 */
    for (int64_t r = d->block_off[b]; r < d->block_off[b + 1]; r++) {
      const int uid = d->run_uid[r];
      float* theta = m->theta + (size_t)uid * m->stride;
      for (int64_t k = d->run_off[r]; k < d->run_off[r + 1]; k++) {
        memset(q, 0, sizeof(float) * (size_t)dim);
        const int vid = d->vid[k];
        float* phi = m->phi + (size_t)vid * m->stride;
        const float rating = d->rating[k];
        memcpy(st->theta_old + (size_t)uid * m->stride, theta, sizeof(float) * (size_t)dim); /* :67 */
        memcpy(st->phi_old + (size_t)vid * m->stride, phi, sizeof(float) * (size_t)dim);     /* :68 */
        float pred = active(o_sdot(dim, theta, phi) + m->bu[uid] + m->bv[vid] + gb, st->loss);
        float error = rating - pred;                                   /* admf.h:70 */
        error = eta * error;                                           /* admf.h:71 */
        o_saxpy(dim, error, theta, q);                                 /* admf.h:72 */
        o_saxpy(dim, -eta * st->lam_u, theta, theta);                  /* admf.h:73 */
        o_saxpy(dim, error, phi, theta);                               /* admf.h:74 */
        o_saxpy(dim, 1.0f - eta * st->lam_v, phi, q);                  /* admf.h:75 */
        memcpy(phi, q, sizeof(float) * (size_t)dim);                   /* admf.h:76 */
        st->bu_old[uid] = m->bu[uid];                                  /* admf.h:77 */
        st->bv_old[vid] = m->bv[vid];                                  /* admf.h:78 */
        m->bu[uid] = (1.0f - eta * st->lam_bu) * m->bu[uid] + error;   /* admf.h:79 */
        m->bv[vid] = (1.0f - eta * st->lam_bv) * m->bv[vid] + error;   /* admf.h:80 */
      }
      /* admf.h:82-83: one validation record per USER, also for users with no records */
      int64_t ii = st->draws ? st->draws[st->draw_pos++] : (int64_t)(rand() % st->nvalid);
      update_reg(m, st, gb, st->val_u[ii], st->val_v[ii], st->val_r[ii]);
    }
  }
  free(q);
}
