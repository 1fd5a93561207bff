/* mf_b200.h - C ABI of the B200 (sm_100a) implementation of experimental-mf's hot path:
 * the blocked SGD matrix-factorization epoch (mf / dpmf / admf) and the test-RMSE pass.
 *
 * This is the drop-in boundary.  The reference has no FFI; its seam is the C++ object model
 * (SURVEY.md 8b): tbb::pipeline calls `void* Filter::operator()(void* block)` on an mf::Block
 * and the filter mutates the public arrays of an MF object.  Each entry point below names the
 * reference interface it replaces (file:line under the reference's src/).  The host-side C++
 * mirror of MF / DPMF / AdaptRegMF (experimental-mf_b200/csrc/model.h) is a thin wrapper over
 * these calls; INTEGRATION.md shows the binding a maintainer of the reference would add.
 *
 * Conventions: plain pointers and sizes only; every call returns 0 on success or a negative
 * MFB_E_* code, with a thread-local message from mfb_last_error().  All factor math is fp32,
 * ids are int32.  A context is bound to one CUDA device and one stream; calls on one context
 * must come from one host thread at a time.  There is no CPU fallback: without a usable CUDA
 * device mfb_create() fails with MFB_E_CUDA.
 */
#ifndef MF_B200_H
#define MF_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mfb_ctx mfb_ctx;

enum {
  MFB_OK = 0,
  MFB_E_ARG = -1,   /* bad argument / state */
  MFB_E_CUDA = -2,  /* CUDA runtime error (message has the cudaError string) */
  MFB_E_NOMEM = -3,
  MFB_E_IO = -4,
  MFB_E_COMM = -5   /* NCCL error */
};

/* model arrays (model.h:23 theta_/phi_/bu_/bv_; model.h:62 ur_/vr_/lambda_u_/lambda_v_;
 * model.h:106 theta_old_/phi_old_/bu_old_/bv_old_) */
enum {
  MFB_THETA = 0, MFB_PHI = 1, MFB_BU = 2, MFB_BV = 3,
  MFB_THETA_OLD = 4, MFB_PHI_OLD = 5, MFB_BU_OLD = 6, MFB_BV_OLD = 7,
  MFB_UR = 8, MFB_VR = 9, MFB_LAMBDA_U = 10, MFB_LAMBDA_V = 11
};

/* update schedules */
enum {
  /* Hogwild: user-runs are handed to sub-warps in file order from a device-side queue; item
   * rows are read and written without synchronisation (the reference with --fly N, mf.h:75). */
  MFB_MODE_HOGWILD = 0,
  /* Ordered: one sub-warp walks the file in file order - the reference's single-thread update
   * order (--fly 1) - using the oracle's operation order (unfused mul/add, index-order dot), so
   * results are bit-comparable with the CPU oracle.  A parity mode, not a fast one. */
  MFB_MODE_ORDERED = 1,
  /* Hogwild schedule, but item rows and item biases are updated with fp32 atomic adds
   * (red.global.add.v4.f32) of the increment, so concurrent updates are never lost. */
  MFB_MODE_ATOMIC = 2
};

const char* mfb_last_error(void);
/* "x.y sm_100a"; also proves the library loads without a GPU */
const char* mfb_version(void);
/* row stride in floats used for a given dim: padding(dim), util.h:163-165 */
int mfb_padding(int dim);

/* ---- context: replaces MF::MF + MF::init storage (model.h:8-14, model.cc:10-21) ------------
 * Allocates theta[nu][stride], phi[nv][stride], bu[nu], bv[nv] in HBM, zero-filled. */
int mfb_create(mfb_ctx** out, int device, int nu, int nv, int dim);
void mfb_destroy(mfb_ctx* ctx);
/* run all work of this context on an existing cudaStream_t (e.g. torch's current stream) */
int mfb_set_stream(mfb_ctx* ctx, void* cuda_stream);
int mfb_sync(mfb_ctx* ctx);
/* tuning knobs (DESIGN.md "concurrency bounds"): "row_concurrency" (default 16: bound on the stale
 * updates of the hottest item row in flight at once, at eta = 0.02), "eta_scaling" (1: that bound
 * widens with 0.02/eta; 2: the run bound too; 0: off), "run_fraction_ppm" (default 3500: user-runs in
 * flight / user-runs of the file; "run_bound_epochs" = n lifts it n epochs after the factors were set - default:
 * never, see mfb_internal.h - and "model_age" is that epoch count, kept by the library, settable by hosts), "max_groups" (explicit number of runs in flight), "kernel"
 * (0 = choose, 1 = generic warp per run, 3 = sub-warp stream, 4 = burst), "ring" (1..4), "throttle",
 * "ctas_per_sm", "threads", "memopt" */
int mfb_set_option(mfb_ctx* ctx, const char* name, int value);
/* allocate the optional array groups: 1 = admf shadows (*_OLD), 2 = dpmf (UR, VR, LAMBDA_*) */
int mfb_enable(mfb_ctx* ctx, int group);

/* ---- factor access: replaces direct theta_[i][j] access (model.h:23, model.cc:85-95) --------
 * Copies rows [row0, row0+nrows) of a model array between a host matrix with `host_stride`
 * floats per row (>= dim; pass 1 for the vector arrays) and the device. */
int mfb_upload(mfb_ctx* ctx, int which, const float* host, int64_t row0, int64_t nrows,
               int64_t host_stride);
int mfb_download(mfb_ctx* ctx, int which, float* host, int64_t row0, int64_t nrows,
                 int64_t host_stride);
/* MF::init's fill (model.cc:22-33): every element ~ N(0,1)*scale from a Philox4x32-10 stream.
 * (The reference uses a clock-seeded engine inside an OpenMP loop, i.e. it is unreproducible;
 * parity tests upload explicit factors instead.) */
int mfb_init_normal(mfb_ctx* ctx, uint64_t seed, float scale);
/* admf init1 (model.cc:369-382): *_OLD <- current */
int mfb_snapshot_old(mfb_ctx* ctx);
/* raw device pointer of a model array (for NCCL / torch interop); NULL if not allocated */
void* mfb_device_ptr(mfb_ctx* ctx, int which);

/* ---- datasets: replace ParseFilter / plain_read (mf.h:57-69, util.h:76-88) ------------------
 * A dataset is a rating file kept in file order as SoA tiles in HBM.  Blocks are appended as
 * flat arrays: users of the block in order (uid[nusers]), rec_off[nusers+1] offsets into
 * vid[]/rating[] relative to the block.  finalize() uploads and builds the schedules. */
int mfb_dataset_create(mfb_ctx* ctx, int* ds);
int mfb_dataset_append_block(mfb_ctx* ctx, int ds, int32_t nusers, const int32_t* uid,
                             const int32_t* rec_off, const int32_t* vid, const float* rating);
/* parse a whole [u32 size][mf.Block] file (getdata.cc:100-103) with the built-in wire decoder */
int mfb_dataset_load_file(mfb_ctx* ctx, int ds, const char* path);
int mfb_dataset_finalize(mfb_ctx* ctx, int ds);
int mfb_dataset_free(mfb_ctx* ctx, int ds);
int64_t mfb_dataset_num_ratings(mfb_ctx* ctx, int ds);
int64_t mfb_dataset_num_runs(mfb_ctx* ctx, int ds);

/* ---- host-side rating files (no GPU needed): replace plain_read (util.h:76-88) and the
 * getdata writer (getdata.cc:82-126).  An mfb_blocks is a parsed [u32 size][mf.Block] file in file
 * order as flat arrays: block_off[nblocks+1] (first run of each block), run_uid[nruns],
 * run_off[nruns+1] (first record of each run), vid[n], rating[n]. */
typedef struct mfb_blocks mfb_blocks;
int mfb_blocks_read(const char* path, mfb_blocks** out);
int mfb_blocks_from_arrays(int64_t nblocks, const int64_t* block_off, int64_t nruns,
                           const int32_t* run_uid, const int32_t* run_off, const int32_t* vid,
                           const float* rating, mfb_blocks** out);
int mfb_blocks_write(const mfb_blocks* b, const char* path);
void mfb_blocks_free(mfb_blocks* b);
int64_t mfb_blocks_num_blocks(const mfb_blocks* b);
int64_t mfb_blocks_num_runs(const mfb_blocks* b);
int64_t mfb_blocks_num_ratings(const mfb_blocks* b);
const int64_t* mfb_blocks_block_off(const mfb_blocks* b);
const int32_t* mfb_blocks_run_uid(const mfb_blocks* b);
const int32_t* mfb_blocks_run_off(const mfb_blocks* b);
const int32_t* mfb_blocks_vid(const mfb_blocks* b);
const float* mfb_blocks_rating(const mfb_blocks* b);
/* append every block of a parsed file to a dataset (before finalize) */
int mfb_dataset_append_blocks(mfb_ctx* ctx, int ds, const mfb_blocks* b);

/* ---- synthetic ratings of a named shape (the reference ships no data).
 * Counter-based (Philox4x32-10), so the result depends only on the parameters - not on the
 * thread count, and a user range [user_begin, user_end) yields exactly that slice of the full
 * data set (used to shard users over GPUs).  Planted rank-`rank` model U*,V* ~ N(0,1/4),
 * r = clamp(round(gb + <u*,v*> + N(0, noise_sd^2)), 1, 5); user degree ~ lognormal(sigma) scaled
 * to nnz; item popularity ~ Zipf(zipf_s) over a random item permutation; no duplicate (u,i).
 * Layout mirrors getdata.cc --method userwise --split S + --method protobuf --size B
 * (getdata.cc:21-126): train records are dealt into `split` chunks, each chunk grouped by user
 * in a random user order, `users_per_block` users per Block.  test/valid are single chunks. */
typedef struct {
  int32_t nu, nv;
  int64_t nnz;             /* total ratings (train + test + valid), approximate */
  int32_t rank;            /* 16 */
  float gb, noise_sd;      /* 2.76, 0.5 */
  float degree_sigma;      /* 1.0 */
  float zipf_s;            /* 1.0 */
  float test_frac, valid_frac;
  int32_t split;           /* 4 */
  int32_t users_per_block; /* 500 */
  uint64_t seed;           /* 0x4D46B200 */
  int32_t user_begin, user_end; /* 0, nu for everything */
  int32_t threads;         /* 0 = hardware concurrency */
} mfb_gen_params;
void mfb_gen_defaults(mfb_gen_params* p, int32_t nu, int32_t nv, int64_t nnz);
int mfb_generate(const mfb_gen_params* p, mfb_blocks** train, mfb_blocks** test,
                 mfb_blocks** valid);

/* ---- the hot path ---------------------------------------------------------------------------
 * SgdFilter::operator() over every block of the file (mf.h:76-132): one SGD epoch with
 * learning rate eta, regulariser lambda, global bias gb. */
int mfb_sgd_epoch(mfb_ctx* ctx, int ds, float eta, float lambda, float gb, int mode);
/* SgdFilter::operator() over Blocks [block_begin, block_end) of the file only (the reference calls the
 * filter once per Block, mf.h:76): a slice of an epoch.  Used by hosts that interleave several files. */
int mfb_sgd_epoch_blocks(mfb_ctx* ctx, int ds, int64_t block_begin, int64_t block_end, float eta, float lambda,
                         float gb, int mode);
int64_t mfb_dataset_num_blocks(mfb_ctx* ctx, int ds);
/* The same epoch with the rating tiles starting in HOST memory (the reference re-reads its
 * training file every epoch, mf.h:24-45): the arrays of `src` (pin them once with
 * mfb_blocks_pin) are copied H2D in chunks of about `chunk_ratings` records (0 = default) on a
 * second stream, overlapped with the update kernel of the previous chunk.  `ds` must be a
 * finalized dataset created from the same blocks; its HBM tiles are the copy target. */
int mfb_sgd_epoch_from_host(mfb_ctx* ctx, int ds, const mfb_blocks* src, float eta, float lambda,
                            float gb, int mode, int64_t chunk_ratings);
/* One epoch straight from the [u32 size][mf.Block] training FILE, nothing resident - the reference's own way of
 * running an epoch (read -> parse -> update pipeline, main.cc:45-50, mf.h:24-69).  The file is cut into chunks of
 * whole Blocks (about `tile_ratings` records, 0 = 8 Mi).  Per chunk: the host cores pread the frames into pinned
 * memory and find the byte range of every serialized mf.User (one jump per user); the RAW bytes go to the GPU, which
 * decodes the varints / records there (mfb_wire_decode.cu; any valid encoding of blocks.proto, not just the canonical
 * one) and runs the update kernel on the decoded tiles while the next chunk is copied and the one after that is read.
 * Three slots of device buffers (raw bytes + 8 bytes per record) whatever the file size; they stay allocated in the
 * context between epochs.  Option "file_decode" = 0 decodes on the host cores instead (pinned SoA chunks).
 * File order is kept: in the ORDERED schedule the result equals mfb_sgd_epoch on the loaded file bit for bit.
 * Errors: MFB_E_IO (unreadable / truncated / malformed), MFB_E_ARG (uid or vid out of range); chunks before the bad
 * one have been applied.  *ratings (optional) = records processed. */
int mfb_sgd_epoch_from_file(mfb_ctx* ctx, const char* path, float eta, float lambda, float gb, int mode,
                            int64_t tile_ratings, int64_t* ratings);
/* Streaming ingest: ONE pass over the training file that leaves the new dataset `ds` finalized in HBM -
 * what mfb_dataset_load_file + mfb_dataset_finalize do, with the records decoded on the GPU - and, with_epoch != 0, runs
 * the SGD epoch on every chunk as it lands: the first epoch costs what an out-of-core epoch costs and the one-time parse
 * + upload stall is gone (the reference overlaps read, parse and update the same way, main.cc:45-50).  ORDERED mode:
 * bit-exact with load + finalize + mfb_sgd_epoch.  Not for contexts with MFB_ENABLE dpmf arrays (logical clock). */
int mfb_dataset_ingest_file(mfb_ctx* ctx, int ds, const char* path, int with_epoch, float eta, float lambda, float gb,
                            int mode, int64_t tile_ratings, int64_t* ratings);
/* the host half of that path alone (no GPU): frames of the file, serialized mf.User messages in them (top-level walk of
 * every Block, blocks.proto:14-16) and the bytes inside those messages; MFB_E_IO on a truncated frame / malformed Block */
int mfb_wire_index_file(const char* path, int64_t* nframes, int64_t* nusers, int64_t* user_bytes);
/* re-send the tiles of finalized dataset `ds` from the (pinned) arrays of `src` on the copy stream;
 * the next epoch kernel on `ds` waits for the copy.  Used by the multi-GPU end-to-end path. */
int mfb_dataset_refresh_from_host(mfb_ctx* ctx, int ds, const mfb_blocks* src);
/* pin: registers the arrays with CUDA and, when the data allow (item ids < 2^24, at most 256 distinct
 * rating values), builds a lossless compact copy (u16 id + u8 code, 3 bytes per record; a second u8 plane with
 * bits 16..23 of the id when some id is >= 65536: 4 bytes) that the streamed epoch sends instead of the 8-byte
 * records and expands on the device */
int mfb_blocks_pin(mfb_blocks* b);
int mfb_blocks_unpin(mfb_blocks* b);
/* MF::calc_mse (model.cc:41-73): SUM of squared errors and the record count. */
int mfb_sse(mfb_ctx* ctx, int ds, float gb, double* sse, int64_t* n);
/* the same pass with the prediction sent through the link of --loss first: link 0 = identity (mfb_sse), 1 =
 * logistic, active() of util.h:90-95 as the update (admf.h:69) and updateReg (model.h:87) apply it.  The reference's
 * calc_mse ignores loss_ (model.cc:62) and never reads measure_ (model.h:109); `--measure 1` selects this here. */
int mfb_sse_link(mfb_ctx* ctx, int ds, float gb, int link, double* sse, int64_t* n);
/* MF::seteta (model.cc:36-38) / DPMF::seteta_cutoff (model.cc:350-352): same formula, host side */
float mfb_seteta(float eta0, int round, float gam);
float mfb_seteta_cutoff(float eta0, int round, float gam, float mineta);

/* ---- dpmf: SGLD / differentially private MF ----------------------------------------------------
 * Call mfb_enable(ctx, 2) BEFORE finalizing the training dataset: finalize then also builds the
 * static logical clock (the values dpmf.h:61-66's counters take in file order).
 * mfb_dp_weights: DPMF::sample_train_and_precompute_weight (model.cc:263-297): fills UR/VR with
 * ntrain/count(u), ntrain/count(v) and returns ntrain.  mfb_dp_bound: model.cc:239-242. */
int mfb_dp_weights(mfb_ctx* ctx, int ds, int32_t* ntrain);
float mfb_dp_bound(float epsilon, int tau, int nv);
typedef struct {
  float eta, temp, bound;          /* eta_, temp_, bound_ (model.h:66-68) */
  int32_t ntrain;
  float lambda_r, lambda_ub, lambda_vb; /* model.h:68; lambda_u_/lambda_v_ are the LAMBDA_U/V arrays */
  uint64_t seed;                   /* Philox key */
  uint32_t round;                  /* epoch number, part of the Philox counter */
  int32_t use_table;               /* ordered parity mode: read noise from the uploaded table ... */
  int32_t table_offset;            /* ... at this fixed offset (the reference's thetaind/phiind) */
} mfb_sgld_params;
/* SgldFilter::operator() over every block (dpmf.h:41-91); mode HOGWILD or ORDERED */
int mfb_sgld_epoch(mfb_ctx* ctx, int ds, const mfb_sgld_params* p, float gb, int mode);
/* DPMF::finish_noise (model.cc:312-332) for the epoch just run on dataset ds */
int mfb_sgld_flush_noise(mfb_ctx* ctx, int ds, const mfb_sgld_params* p);
/* the reductions DPMF::sample_hyper needs (model.cc:336-342): column sums of squares of theta
 * and phi (normu[dim], normv[dim]) and |bu|^2, |bv|^2, accumulated in fp64 */
int mfb_col_sqnorms(mfb_ctx* ctx, double* normu, double* normv, double* bu2, double* bv2);
/* upload a stand-in for the reference's noise_ lookup table (model.cc:229-231); parity tests only */
int mfb_set_noise_table(mfb_ctx* ctx, const float* host, int64_t n);

/* ---- admf: adaptive regulariser (Rendle) ------------------------------------------------------
 * Needs mfb_enable(ctx, 1) (the *_OLD shadows; mfb_snapshot_old() is AdaptRegMF::init1's copy).
 * set_validation: the flattened, already shuffled validation list recsv_ (model.cc:390-415).
 * set_draws: for the next epoch, one index into that list per user-run in file order - the host
 * draws them with rand() % |valid| exactly as admf.h:82 does, so the sequence is the reference's.
 * lams: lam_u_, lam_v_, lam_bu_, lam_bv_ (model.h:111-117), all initialised to --lambda. */
int mfb_admf_set_validation(mfb_ctx* ctx, int64_t n, const int32_t* u, const int32_t* v, const float* r);
int mfb_admf_set_draws(mfb_ctx* ctx, int64_t n, const int32_t* draws);
int mfb_admf_set_lams(mfb_ctx* ctx, const float lams[4]);
int mfb_admf_get_lams(mfb_ctx* ctx, float lams[4]);
/* AdRegFilter::operator() over every block (admf.h:52-86) + updateReg after every user
 * (model.h:86-102); eta_reg = AdaptRegMF::set_etareg's value (model.cc:386-388);
 * mode ORDERED (serial, exact lambda trajectory) or ATOMIC (parallel) */
int mfb_admf_epoch(mfb_ctx* ctx, int ds, float eta, float eta_reg, int loss, float gb, int mode);

/* ---- multi-GPU: DSGD stratification (new design; the reference is single-process) -------------
 * One process per GPU.  Users are sharded over ranks, items are cut into `world` blocks by
 * item_bounds[world+1]; datasets[j] holds this rank's ratings whose item is in block j (see
 * mfb_blocks_split_by_item).  mfb_dsgd_epoch runs `world` sub-epochs; after each one the item
 * block just updated is passed to rank-1 and the next one received from rank+1 with
 * ncclSend/ncclRecv on a dedicated stream.  At entry and exit rank r holds block r.
 * The unique id is created on rank 0 and distributed by the host program (bench.py uses
 * torch.distributed for that plumbing). */
int mfb_blocks_split_by_item(const mfb_blocks* b, int nparts, const int32_t* bounds, mfb_blocks** out);
/* one run per user (its runs concatenated in file order, users in order of first appearance) */
int mfb_blocks_merge_runs(const mfb_blocks* b, int users_per_block, mfb_blocks** out);
/* the general regrouping: merge_users != 0 as above; longest_first = 1: runs in descending order of length (a
 * launch ends when its longest run still in flight ends: longest-processing-time-first scheduling); longest_first =
 * N > 1: only the runs of N records or more move to the front, the others keep their order.  Changes the
 * order of updates: for epochs after the first, where the order no longer shows in the result. */
int mfb_blocks_regroup(const mfb_blocks* b, int merge_users, int longest_first, int users_per_block, mfb_blocks** out);
int mfb_comm_unique_id(void* out128);
int mfb_comm_init(mfb_ctx* ctx, int rank, int world, const void* id128);
int mfb_comm_destroy(mfb_ctx* ctx);
/* Peer-memory ring (optional; without it the shifts are ncclSend/ncclRecv).  export: three CUDA IPC handles (3 x 64
 * bytes) of the allocations holding this rank's phi, bv and arrival flags plus the two byte offsets of phi and bv
 * inside them (208 bytes in all); the host program hands them to rank+1 (any transport), which calls
 * import with the handles of the rank it sends to (rank-1).  From then on a shift is one copy per array straight into
 * the neighbour's HBM over NVLink plus a sequence number the neighbour's compute stream waits for on the device.
 * export fails while the placement search may still relocate the item matrix. */
int mfb_comm_ipc_export(mfb_ctx* ctx, void* out208);
int mfb_comm_ipc_import(mfb_ctx* ctx, const void* in208);
/* unmap the neighbour's memory again; call on every rank and synchronise the ranks BEFORE any of them destroys its
 * context (CUDA: an importer must close its mapping before the exporter frees the memory) */
int mfb_comm_ipc_close(mfb_ctx* ctx);
int mfb_dsgd_epoch(mfb_ctx* ctx, const int* datasets, const int32_t* item_bounds, float eta, float lambda,
                   float gb, int mode);
/* The general form.  halves = H >= 1: every rank's item block is cut into H pieces (item_bounds has world*H+1
 * entries, datasets world*H cells, block of rank r = pieces r*H .. r*H+H-1); the shift of piece h travels on the
 * communication stream while the kernel on piece h+1 runs, so the exchange and the ring neighbour's lag are hidden
 * behind compute (H = 2: the overlap SURVEY.md 8e asks for).  rotations = R >= 1: the ring turns R times in the
 * epoch, turn r over the r-th of R runs of consecutive Blocks of every cell.  R > 1 in the FIRST epoch keeps the
 * multi-GPU test RMSE on the single-GPU trajectory; later epochs use R = 1.
 * mfb_dsgd_epoch == halves 1, rotations 1. */
int mfb_dsgd_epoch_ex(mfb_ctx* ctx, const int* datasets, const int32_t* item_bounds, int halves, int rotations,
                      float eta, float lambda, float gb, int mode);
/* diagnostic: the most recent DSGD epoch on this rank, kernel by kernel in launch order: out[2i] = ms the compute
 * stream waited before kernel i for its item rows (ring shift issued one sub-epoch earlier), out[2i+1] = ms of
 * kernel i; returns the number of entries written (<= n) or a negative error */
int mfb_dsgd_timeline(mfb_ctx* ctx, float* out, int n);
/* make all of phi/bv valid on every rank (each rank publishes its home block) - before evaluation */
int mfb_comm_allgather_items(mfb_ctx* ctx, const int32_t* item_bounds);
/* sum (sse, n) over the ranks */
int mfb_comm_allreduce_sse(mfb_ctx* ctx, double* sse, int64_t* n);

/* Staleness probe of the parallel SGD schedule (diagnostic): with the probe armed (item >= 0; item < 0
 * disarms) every update also counts, per item, the updates the L2 has performed, and records how
 * many updates of the same item happened between the read of its row and the update.
 * mfb_probe_read: out = {stale updates summed over all updates, number of updates, the same two
 * for the probed item}; the counters restart. */
int mfb_probe_arm(mfb_ctx* ctx, int item);
int mfb_probe_read(mfb_ctx* ctx, uint64_t out[4]);

/* device time in ms of the most recent epoch / sse call's kernels (CUDA events on the
 * context's stream; valid after mfb_sync) and the number of kernel launches since create */
float mfb_last_kernel_ms(mfb_ctx* ctx);
int64_t mfb_launch_count(mfb_ctx* ctx);
/* bytes copied host -> device by mfb_sgd_epoch_from_host since the context was created */
int64_t mfb_h2d_bytes(mfb_ctx* ctx);
/* Placement search of the item matrix: before the first parallel SGD epoch on a file of
 * at least "placement_min_ratings" records (option, default 4,000,000) the library tries
 * "placement_trials" (option, default 16; <= 1 turns it off) locations and keeps the fastest - for the
 * plane-layout working copy of phi that whole-epoch launches use (which = 1) and, on the first chunked or
 * multi-GPU epoch, for the rows of phi themselves (which = 0); bv moves with the first search.
 * mfb_device_ptr(MFB_PHI / MFB_BV) may therefore change at those epochs.  The report: ms[i] =
 * calibration time of candidate i, *best = the one kept; returns the number of candidates tried. */
int mfb_placement_report(mfb_ctx* ctx, int which, float* ms, int n, int* best);
/* shape of the most recent SGD epoch launch: out = {kernel variant, grid, threads per CTA, ring depth} */
int mfb_last_launch(mfb_ctx* ctx, int out[4]);

#ifdef __cplusplus
}
#endif
#endif
