// C1: DSGD stratification over the GPUs of one NVSwitch box.  New design - the
// reference has no multi-process path.
//
// One process per GPU.  Users are sharded by rank: rank p owns theta/bu of its users and all of
// their ratings, pre-split by item block.  Items are cut into P blocks.  An epoch is P
// sub-epochs: in sub-epoch s rank p updates cell (p, b = (p+s) mod P) with block b of phi/bv
// resident, then hands block b to rank p-1 and receives block b+1 from rank p+1 (a ring shift over
// NVLink, ncclSend/ncclRecv inside one group on a dedicated stream).  Cells that run at the same
// time share neither users nor items, so there is no inter-GPU conflict; inside a cell the
// schedule is the single-GPU one.  After P shifts every block is home again.
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <vector>

#include "mfb_internal.h"

#define MFB_MAX_HALVES 4

namespace mfb {

// NCCL is bound at run time (dlopen), not at link time: the library then loads on machines
// without NCCL, and inside a PyTorch process it shares the libnccl.so.2 torch already mapped
// instead of pulling a second copy with the same SONAME.  MFB_NCCL_LIB overrides the path.
struct NcclApi {
  void* handle = nullptr;
  decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
  decltype(&ncclCommInitRank) CommInitRank = nullptr;
  decltype(&ncclCommDestroy) CommDestroy = nullptr;
  decltype(&ncclGroupStart) GroupStart = nullptr;
  decltype(&ncclGroupEnd) GroupEnd = nullptr;
  decltype(&ncclSend) Send = nullptr;
  decltype(&ncclRecv) Recv = nullptr;
  decltype(&ncclBroadcast) Broadcast = nullptr;
  decltype(&ncclAllReduce) AllReduce = nullptr;
  decltype(&ncclGetErrorString) GetErrorString = nullptr;
};
static NcclApi g_nccl;

static int load_nccl() {
  if (g_nccl.handle) return MFB_OK;
  const char* path = getenv("MFB_NCCL_LIB");
  void* h = dlopen(path ? path : "libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) {
    set_error("cannot load NCCL: %s", dlerror());
    return MFB_E_COMM;
  }
#define MFB_SYM(name)                                                   \
  g_nccl.name = (decltype(g_nccl.name))dlsym(h, "nccl" #name);          \
  if (!g_nccl.name) {                                                   \
    set_error("NCCL symbol nccl" #name " missing");                     \
    return MFB_E_COMM;                                                  \
  }
  MFB_SYM(GetUniqueId) MFB_SYM(CommInitRank) MFB_SYM(CommDestroy) MFB_SYM(GroupStart) MFB_SYM(GroupEnd)
  MFB_SYM(Send) MFB_SYM(Recv) MFB_SYM(Broadcast) MFB_SYM(AllReduce) MFB_SYM(GetErrorString)
#undef MFB_SYM
  g_nccl.handle = h;
  return MFB_OK;
}

#define MFB_NCCL(expr)                                                                    \
  do {                                                                                    \
    ncclResult_t _r = (expr);                                                             \
    if (_r != ncclSuccess) {                                                              \
      set_error("%s:%d %s: %s", __FILE__, __LINE__, #expr, g_nccl.GetErrorString(_r));       \
      return MFB_E_COMM;                                                                  \
    }                                                                                     \
  } while (0)

struct Comm {
  ncclComm_t nccl = nullptr;
  int rank = 0, world = 1;
  cudaStream_t stream = nullptr;
  cudaEvent_t computed[MFB_MAX_HALVES] = {nullptr}, shifted[MFB_MAX_HALVES] = {nullptr};
  bool shift_pending[MFB_MAX_HALVES] = {false};
  double* d_red = nullptr;  // [2] sse, count
  // Peer-memory ring (mfb_comm_ipc_export / import): the shift PUSHES the block into the neighbour's copy of phi / bv
  // over NVLink (its allocations are mapped here through CUDA IPC) and then raises a sequence number in the
  // neighbour's flag word; the neighbour's compute stream waits for that number on the device.  No NCCL call, no
  // host involvement, one copy per array.
  float* peer_phi = nullptr;   // rank-1's phi / bv / flags, mapped
  float* peer_bv = nullptr;
  void *map_phi = nullptr, *map_bv = nullptr;  // what cudaIpcOpenMemHandle returned (allocation bases)
  unsigned* peer_flags = nullptr;
  unsigned* d_flags = nullptr;  // [MFB_MAX_HALVES + 1]: [slot] = shifts of that piece slot that have arrived here; [last] = timeout seen
  unsigned seq[MFB_MAX_HALVES] = {0};  // shifts issued per piece slot (the same count on both ends of a link)
  unsigned* h_err = nullptr;    // pinned copy of the timeout word
  bool p2p = false;
  // diagnostic timeline of the most recent epoch (mfb_dsgd_timeline): events on the compute stream
  // at the start of the epoch, after every cell kernel and after every wait for the ring shift
  std::vector<cudaEvent_t> marks;
  int nmarks = 0;
};

static void close_peer(Comm* m) {
  if (m->map_phi) cudaIpcCloseMemHandle(m->map_phi);
  if (m->map_bv) cudaIpcCloseMemHandle(m->map_bv);
  if (m->peer_flags) cudaIpcCloseMemHandle(m->peer_flags);
  m->map_phi = m->map_bv = nullptr;
  m->peer_phi = m->peer_bv = nullptr;
  m->peer_flags = nullptr;
  m->p2p = false;
}

// base and size of the cudaMalloc allocation that holds p
static bool cuda_range(const void* p, void** base, size_t* size) {
  typedef int (*GetRange)(unsigned long long*, size_t*, unsigned long long);
  static GetRange fn = (GetRange)dlsym(RTLD_DEFAULT, "cuMemGetAddressRange_v2");
  if (!fn) {
    void* h = dlopen("libcuda.so.1", RTLD_NOW | RTLD_GLOBAL);
    if (h) fn = (GetRange)dlsym(h, "cuMemGetAddressRange_v2");
  }
  unsigned long long b = 0;
  if (!fn || fn(&b, size, (unsigned long long)(uintptr_t)p) != 0) return false;
  *base = (void*)(uintptr_t)b;
  return true;
}

// the sequence number of a delivered shift, visible to the neighbour after everything this stream copied before it
__global__ void ring_flag_set_kernel(unsigned* peer_flag, unsigned seq) {
  __threadfence_system();
  *reinterpret_cast<volatile unsigned*>(peer_flag) = seq;
  __threadfence_system();
}
// wait (on the device, one thread) until the neighbour has delivered shift `seq`; gives up after `timeout_ns` and
// says so in *err - a lost neighbour must not hang the GPU
__global__ void ring_flag_wait_kernel(const unsigned* flag, unsigned seq, unsigned long long timeout_ns, unsigned* err) {
  unsigned long long t0, t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  while ((int)(*reinterpret_cast<const volatile unsigned*>(flag) - seq) < 0) {
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    if (t - t0 > timeout_ns) {
      *err = seq ? seq : 1u;
      break;
    }
    __nanosleep(200);
  }
  __threadfence_system();
}

}  // namespace mfb

using namespace mfb;

extern "C" {

// ---- peer-memory ring --------------------------------------------------------------------------------------------
// out208 = three cudaIpcMemHandle_t (64 bytes each) - the allocations that hold this rank's phi, bv and flag words -
// followed by two uint64: the byte offsets of phi and bv inside their allocations (after the placement search the two
// live inside its arena).  Fails (MFB_E_ARG) while the placement search may still relocate the item matrix (matrices
// that fit the L2, before their first parallel epoch): the mapping would go stale.
int mfb_comm_ipc_export(mfb_ctx* h, void* out208) {
  MFB_REQUIRE(h && out208, "NULL argument");
  Context* c = &h->c;
  Comm* m = (Comm*)c->comm;
  MFB_REQUIRE(m, "communicator not initialised");
  const size_t phi_bytes = (size_t)c->nv * c->stride * sizeof(float);
  MFB_REQUIRE(c->opt_placement_trials <= 1 || phi_bytes > ((size_t)64 << 20) || c->placement_done[0],
              "the placement search may still move the item matrix: run one epoch first");
  MFB_CUDA(cudaSetDevice(c->device));
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  if (!m->d_flags) {
    MFB_CUDA(cudaMalloc(&m->d_flags, (MFB_MAX_HALVES + 1) * sizeof(unsigned)));
    MFB_CUDA(cudaMemset(m->d_flags, 0, (MFB_MAX_HALVES + 1) * sizeof(unsigned)));
    MFB_CUDA(cudaMallocHost(&m->h_err, sizeof(unsigned)));
    *m->h_err = 0;
  }
  auto base_of = [&](const float* p, uint64_t* off) -> const void* {  // the allocation a pointer lives in
    const char* q = (const char*)p;
    if (c->placement_arena && q >= c->placement_arena) {
      // (slots of the arena are laid out behind each other; the arena is one cudaMalloc)
      size_t asize = 0;
      void* abase = nullptr;
      if (cuda_range(c->placement_arena, &abase, &asize) && q < (const char*)abase + asize) {
        *off = (uint64_t)(q - (const char*)abase);
        return abase;
      }
    }
    *off = 0;
    return p;
  };
  cudaIpcMemHandle_t* out = (cudaIpcMemHandle_t*)out208;
  uint64_t off[2];
  const void* bphi = base_of(c->arr[MFB_PHI], &off[0]);
  const void* bbv = base_of(c->arr[MFB_BV], &off[1]);
  MFB_CUDA(cudaIpcGetMemHandle(&out[0], (void*)bphi));
  MFB_CUDA(cudaIpcGetMemHandle(&out[1], (void*)bbv));
  MFB_CUDA(cudaIpcGetMemHandle(&out[2], m->d_flags));
  memcpy((char*)out208 + 192, off, sizeof off);
  return MFB_OK;
}

// Unmaps the neighbour's memory (the ring falls back to ncclSend/ncclRecv).  CUDA wants every importer to have closed
// its mapping before the exporter frees the memory: the host program calls this on all ranks, synchronises them, and
// only then destroys the contexts (mfb_dsgd.DsgdWorker.close).
int mfb_comm_ipc_close(mfb_ctx* h) {
  MFB_REQUIRE(h, "ctx is NULL");
  Context* c = &h->c;
  Comm* m = (Comm*)c->comm;
  if (!m) return MFB_OK;
  MFB_CUDA(cudaSetDevice(c->device));
  MFB_CUDA(cudaStreamSynchronize(m->stream));  // this rank's last pushes
  MFB_CUDA(cudaStreamSynchronize(c->stream));
  close_peer(m);
  return MFB_OK;
}

// in208 = what rank-1 (the rank this one sends to) exported.  From here on the ring shifts go through peer memory.
int mfb_comm_ipc_import(mfb_ctx* h, const void* in208) {
  MFB_REQUIRE(h && in208, "NULL argument");
  Context* c = &h->c;
  Comm* m = (Comm*)c->comm;
  MFB_REQUIRE(m && m->d_flags, "export this rank's handles first");
  MFB_REQUIRE(!m->p2p, "peer memory already mapped");
  MFB_CUDA(cudaSetDevice(c->device));
  const cudaIpcMemHandle_t* in = (const cudaIpcMemHandle_t*)in208;
  uint64_t off[2];
  memcpy(off, (const char*)in208 + 192, sizeof off);
  MFB_CUDA(cudaIpcOpenMemHandle(&m->map_phi, in[0], cudaIpcMemLazyEnablePeerAccess));
  if (memcmp(&in[0], &in[1], sizeof in[0]) == 0) {  // phi and bv in one allocation (the arena): one mapping
    m->map_bv = nullptr;
    m->peer_bv = (float*)((char*)m->map_phi + off[1]);
  } else {
    MFB_CUDA(cudaIpcOpenMemHandle(&m->map_bv, in[1], cudaIpcMemLazyEnablePeerAccess));
    m->peer_bv = (float*)((char*)m->map_bv + off[1]);
  }
  m->peer_phi = (float*)((char*)m->map_phi + off[0]);
  MFB_CUDA(cudaIpcOpenMemHandle((void**)&m->peer_flags, in[2], cudaIpcMemLazyEnablePeerAccess));
  m->p2p = true;
  return MFB_OK;
}

int mfb_comm_unique_id(void* out128) {
  MFB_REQUIRE(out128, "NULL argument");
  if (int rc = load_nccl()) return rc;
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  MFB_NCCL(g_nccl.GetUniqueId((ncclUniqueId*)out128));
  return MFB_OK;
}

int mfb_comm_init(mfb_ctx* h, int rank, int world, const void* id128) {
  MFB_REQUIRE(h && id128 && world >= 1 && rank >= 0 && rank < world, "bad argument");
  Context* c = &h->c;
  MFB_REQUIRE(!c->comm, "communicator already initialised");
  if (int rc = load_nccl()) return rc;
  MFB_CUDA(cudaSetDevice(c->device));
  Comm* m = new Comm();
  m->rank = rank;
  m->world = world;
  ncclUniqueId id;
  memcpy(&id, id128, sizeof id);
  MFB_NCCL(g_nccl.CommInitRank(&m->nccl, world, id, rank));
  MFB_CUDA(cudaStreamCreateWithFlags(&m->stream, cudaStreamNonBlocking));
  for (int i = 0; i < MFB_MAX_HALVES; i++) {
    MFB_CUDA(cudaEventCreateWithFlags(&m->computed[i], cudaEventDisableTiming));
    MFB_CUDA(cudaEventCreateWithFlags(&m->shifted[i], cudaEventDisableTiming));
  }
  MFB_CUDA(cudaMalloc(&m->d_red, 2 * sizeof(double)));
  c->comm = m;
  return MFB_OK;
}

int mfb_comm_destroy(mfb_ctx* h) {
  MFB_REQUIRE(h, "ctx is NULL");
  Context* c = &h->c;
  Comm* m = (Comm*)c->comm;
  if (!m) return MFB_OK;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(m->stream);
  cudaStreamSynchronize(c->stream);
  close_peer(m);
  cudaFree(m->d_flags);
  cudaFreeHost(m->h_err);
  if (m->nccl) g_nccl.CommDestroy(m->nccl);
  for (cudaEvent_t e : m->marks) cudaEventDestroy(e);
  for (int i = 0; i < MFB_MAX_HALVES; i++) {
    cudaEventDestroy(m->computed[i]);
    cudaEventDestroy(m->shifted[i]);
  }
  cudaStreamDestroy(m->stream);
  cudaFree(m->d_red);
  delete m;
  c->comm = nullptr;
  return MFB_OK;
}

// Ring shift of one item block, asynchronous to the compute stream: after the compute stream reached the
// point where this is called (the cell kernel on block `b` has been queued), the comm stream sends rows
// [bounds[b], bounds[b+1]) of phi/bv to rank-1 and receives block `nb` from rank+1.  The compute stream does
// NOT wait here: the kernel that next touches block nb waits for m->shifted[slot] (wait_shift).
static int shift_block(Context* c, Comm* m, const int32_t* bounds, int b, int nb, int slot) {
  const int P = m->world;
  const int to = (m->rank + P - 1) % P, from = (m->rank + 1) % P;
  const int64_t s0 = bounds[b], s1 = bounds[b + 1], r0 = bounds[nb], r1 = bounds[nb + 1];
  MFB_CUDA(cudaEventRecord(m->computed[slot], c->stream));
  MFB_CUDA(cudaStreamWaitEvent(m->stream, m->computed[slot], 0));
  if (m->p2p && c->opt_ring_peer) {
    // push: the rows this rank has just updated go straight into rank-1's arrays (it is not touching that block: it
    // works on another one and shipped its previous copy of this one P-1 sub-epochs ago), then the sequence number
    MFB_CUDA(cudaMemcpyAsync(m->peer_phi + s0 * c->stride, c->arr[MFB_PHI] + s0 * c->stride,
                             (size_t)(s1 - s0) * c->stride * sizeof(float), cudaMemcpyDefault, m->stream));
    MFB_CUDA(cudaMemcpyAsync(m->peer_bv + s0, c->arr[MFB_BV] + s0, (size_t)(s1 - s0) * sizeof(float), cudaMemcpyDefault,
                             m->stream));
    ring_flag_set_kernel<<<1, 1, 0, m->stream>>>(m->peer_flags + slot, ++m->seq[slot]);
    MFB_CUDA(cudaGetLastError());
    m->shift_pending[slot] = true;
    (void)r0; (void)r1; (void)from;
    return MFB_OK;
  }
  MFB_NCCL(g_nccl.GroupStart());
  MFB_NCCL(g_nccl.Send(c->arr[MFB_PHI] + s0 * c->stride, (size_t)(s1 - s0) * c->stride, ncclFloat, to, m->nccl, m->stream));
  MFB_NCCL(g_nccl.Send(c->arr[MFB_BV] + s0, (size_t)(s1 - s0), ncclFloat, to, m->nccl, m->stream));
  MFB_NCCL(g_nccl.Recv(c->arr[MFB_PHI] + r0 * c->stride, (size_t)(r1 - r0) * c->stride, ncclFloat, from, m->nccl, m->stream));
  MFB_NCCL(g_nccl.Recv(c->arr[MFB_BV] + r0, (size_t)(r1 - r0), ncclFloat, from, m->nccl, m->stream));
  MFB_NCCL(g_nccl.GroupEnd());
  MFB_CUDA(cudaEventRecord(m->shifted[slot], m->stream));
  m->shift_pending[slot] = true;
  return MFB_OK;
}
static int wait_shift(Context* c, Comm* m, int slot) {
  if (m->shift_pending[slot]) {
    if (m->p2p && c->opt_ring_peer) {  // the block rank+1 pushed here: its sequence number equals the count of this rank's own shifts
      ring_flag_wait_kernel<<<1, 1, 0, c->stream>>>(m->d_flags + slot, m->seq[slot], 10000000000ull, m->d_flags + MFB_MAX_HALVES);
      MFB_CUDA(cudaGetLastError());
    } else {
      MFB_CUDA(cudaStreamWaitEvent(c->stream, m->shifted[slot], 0));
    }
  }
  m->shift_pending[slot] = false;
  return MFB_OK;
}

// One DSGD epoch.  Items are cut into P*H blocks (H = `halves` pieces of each rank's home block); datasets[j]
// holds this rank's ratings with an item in block j.  The epoch is `rotations` turns of the ring; turn r works on
// slice r of every cell (the source file cut into `rotations` runs of consecutive Blocks - a cell keeps the Block
// structure of the file it was split from).  Sub-epoch s of a turn: for h = 0..H-1 the kernel on piece h of block
// (rank+s) mod P, then - on the comm stream, overlapped with the kernel on piece h+1 - the shift of that piece.
//  * H = 2 (default of the host programs) hides the exchange and the neighbour's lag of up to half a sub-epoch
//    behind compute (SURVEY 8e "splitting each item block in two halves");
//  * rotations > 1 in the FIRST epoch keeps the multi-GPU result on the reference's trajectory: the cost of the
//    DSGD order in test RMSE is paid in epoch 1 only, while the factors leave their random initialisation and
//    every (user shard, item block) pair has to agree on the latent basis; the ring turning 16 times in that
//    epoch mixes the pairs 16 times as often (tools/dsgd_order_study.py: P = 8, final tRMSE +0.0142 with one
//    turn in every epoch, +0.0007 with 16 turns in epoch 1 and one afterwards).
int mfb_dsgd_epoch_ex(mfb_ctx* h, const int* datasets, const int32_t* item_bounds, int halves, int rotations,
                      float eta, float lambda, float gb, int mode) {
  MFB_REQUIRE(h && datasets && item_bounds, "NULL argument");
  Context* c = &h->c;
  Comm* m = (Comm*)c->comm;
  MFB_REQUIRE(m, "communicator not initialised");
  MFB_REQUIRE(mode == MFB_MODE_HOGWILD || mode == MFB_MODE_ATOMIC || mode == MFB_MODE_ORDERED, "bad mode");
  MFB_REQUIRE(halves >= 1 && halves <= MFB_MAX_HALVES, "halves must be 1..%d", MFB_MAX_HALVES);
  MFB_REQUIRE(rotations >= 1 && rotations <= 4096, "rotations must be 1..4096");
  const int P = m->world, H = halves, NB = P * H;
  MFB_REQUIRE(item_bounds[0] == 0 && item_bounds[NB] == c->nv, "item_bounds must span [0, nv] in world*halves blocks");
  MFB_CUDA(cudaSetDevice(c->device));
  if (m->p2p && m->h_err && *m->h_err) {
    set_error("DSGD ring: the block of shift %u never arrived from rank %d (10 s)", *m->h_err, (m->rank + 1) % P);
    return MFB_E_COMM;
  }
  std::vector<Dataset*> cells(NB);
  bool resident = true;
  for (int j = 0; j < NB; j++) {
    const int ds = datasets[j];
    MFB_REQUIRE(ds >= 0 && ds < (int)c->datasets.size() && c->datasets[ds].finalized, "bad dataset for block %d", j);
    cells[j] = &c->datasets[ds];
    resident = resident && !cells[j]->refresh_pending;
  }
  // placement of this rank's copy of the item matrix, over all of its cells
  if (!c->placement_done[0] && resident) {
    if (int trc = tune_placement(c, cells.data(), NB, gb, mode, false)) return trc;
  }
  cudaEventRecord(c->ev0, c->stream);
  const int nmarks = 2 * P * H * rotations + 1;
  if ((int)m->marks.size() < nmarks) {
    const size_t old = m->marks.size();
    m->marks.resize(nmarks);
    for (size_t i = old; i < m->marks.size(); i++) MFB_CUDA(cudaEventCreate(&m->marks[i]));
  }
  m->nmarks = 0;
  MFB_CUDA(cudaEventRecord(m->marks[m->nmarks++], c->stream));
  for (int r = 0; r < rotations; r++) {
    for (int s = 0; s < P; s++) {
      for (int hh = 0; hh < H; hh++) {
        const int b = ((m->rank + s) % P) * H + hh, nb = ((m->rank + s + 1) % P) * H + hh;
        Dataset* d = cells[b];
        // the previous shift of this piece slot delivered the rows this kernel works on
        if (int rc = wait_shift(c, m, hh)) return rc;
        MFB_CUDA(cudaEventRecord(m->marks[m->nmarks++], c->stream));
        const int64_t k0 = d->nblocks * r / rotations, k1 = d->nblocks * (r + 1) / rotations;
        const int64_t r0 = d->h_block_off[k0], r1 = d->h_block_off[k1];
        if (r1 > r0) {
          int rc = launch_sgd(c, d, eta, lambda, gb, mode, r0, r1);
          if (rc) return rc;
        }
        MFB_CUDA(cudaEventRecord(m->marks[m->nmarks++], c->stream));
        if (P > 1) {
          int rc = shift_block(c, m, item_bounds, b, nb, hh);
          if (rc) return rc;
        }
      }
    }
  }
  for (int hh = 0; hh < H; hh++)  // every block is home again before anything else uses the item matrix
    if (int rc = wait_shift(c, m, hh)) return rc;
  if (m->p2p && c->opt_ring_peer)  // (checked at the start of the next epoch and by mfb_comm_allreduce_sse)
    MFB_CUDA(cudaMemcpyAsync(m->h_err, m->d_flags + MFB_MAX_HALVES, sizeof(unsigned), cudaMemcpyDeviceToHost, c->stream));
  cudaEventRecord(c->ev1, c->stream);
  c->timed = true;
  c->model_age++;
  return MFB_OK;
}

int mfb_dsgd_epoch(mfb_ctx* h, const int* datasets, const int32_t* item_bounds, float eta, float lambda,
                   float gb, int mode) {
  return mfb_dsgd_epoch_ex(h, datasets, item_bounds, 1, 1, eta, lambda, gb, mode);
}

// Diagnostic: where the most recent DSGD epoch spent its time on this rank.  Steps i = 0, 1, ... are the
// (turn, sub-epoch, piece) kernels in launch order: out[2i] = ms the compute stream waited before kernel i
// for the arrival of its item rows (the ring shift issued one sub-epoch earlier: the neighbour's kernel on
// that piece + the transfer), out[2i+1] = ms of kernel i.
int mfb_dsgd_timeline(mfb_ctx* h, float* out, int n) {
  MFB_REQUIRE(h && out, "NULL argument");
  Context* c = &h->c;
  Comm* m = (Comm*)c->comm;
  MFB_REQUIRE(m && m->nmarks > 1, "no epoch recorded");
  MFB_CUDA(cudaSetDevice(c->device));
  MFB_CUDA(cudaEventSynchronize(m->marks[m->nmarks - 1]));
  for (int i = 0; i + 1 < m->nmarks && i < n; i++) MFB_CUDA(cudaEventElapsedTime(out + i, m->marks[i], m->marks[i + 1]));
  return std::min(n, m->nmarks - 1);
}

// every rank publishes its home block (block == rank): afterwards all ranks hold all of phi/bv
int mfb_comm_allgather_items(mfb_ctx* h, const int32_t* item_bounds) {
  MFB_REQUIRE(h && item_bounds, "NULL argument");
  Context* c = &h->c;
  Comm* m = (Comm*)c->comm;
  MFB_REQUIRE(m, "communicator not initialised");
  MFB_CUDA(cudaSetDevice(c->device));
  MFB_NCCL(g_nccl.GroupStart());
  for (int r = 0; r < m->world; r++) {
    const int64_t r0 = item_bounds[r], r1 = item_bounds[r + 1];
    float* p = c->arr[MFB_PHI] + r0 * c->stride;
    MFB_NCCL(g_nccl.Broadcast(p, p, (size_t)(r1 - r0) * c->stride, ncclFloat, r, m->nccl, c->stream));
    MFB_NCCL(g_nccl.Broadcast(c->arr[MFB_BV] + r0, c->arr[MFB_BV] + r0, (size_t)(r1 - r0), ncclFloat, r, m->nccl, c->stream));
  }
  MFB_NCCL(g_nccl.GroupEnd());
  return MFB_OK;
}

int mfb_comm_allreduce_sse(mfb_ctx* h, double* sse, int64_t* n) {
  MFB_REQUIRE(h && sse && n, "NULL argument");
  Context* c = &h->c;
  Comm* m = (Comm*)c->comm;
  MFB_REQUIRE(m, "communicator not initialised");
  MFB_CUDA(cudaSetDevice(c->device));
  double hb[2] = {*sse, (double)*n};
  MFB_CUDA(cudaMemcpyAsync(m->d_red, hb, sizeof hb, cudaMemcpyHostToDevice, c->stream));
  MFB_NCCL(g_nccl.AllReduce(m->d_red, m->d_red, 2, ncclDouble, ncclSum, m->nccl, c->stream));
  MFB_CUDA(cudaMemcpyAsync(hb, m->d_red, sizeof hb, cudaMemcpyDeviceToHost, c->stream));
  MFB_CUDA(cudaStreamSynchronize(c->stream));
  if (m->p2p && m->h_err && *m->h_err) {
    set_error("DSGD ring: the block of shift %u never arrived from rank %d (10 s)", *m->h_err, (m->rank + 1) % m->world);
    return MFB_E_COMM;
  }
  *sse = hb[0];
  *n = (int64_t)(hb[1] + 0.5);
  return MFB_OK;
}

}  // extern "C"
