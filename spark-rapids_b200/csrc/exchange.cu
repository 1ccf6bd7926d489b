// exchange.cu — (e): the shuffle exchange as an NCCL all-to-all over NVLink 5 / NVSwitch.
// Replaces the reference's RapidsShuffleManager data path (GpuShuffleExchangeExecBase.scala:384-536;
// RapidsShuffleInternalManagerBase.scala:1618 RapidsCachingWriter, :1978 getReaderImpl; the UCX
// transport in shuffle-plugin/): the reference copies each partitioned batch D2H, serialises it on
// CPU threads and moves it through disk/netty or UCX bounce buffers.  Here the partitioned table
// (hash.cu: contiguous per-destination row ranges) stays in HBM: one ncclAllGather of the size
// matrix, then ONE grouped ncclSend/ncclRecv launch moves every column slice peer to peer.
//
// NCCL is bound at run time (dlopen libnccl.so.2) so that the process shares the NCCL that
// torch.distributed already loaded; rendezvous (the 128-byte unique id) is the caller's job.
#include <dlfcn.h>
#include <nccl.h>
#include <mutex>
#include "prim.cuh"

namespace b2 {

struct NcclApi {
  ncclResult_t (*GetUniqueId)(ncclUniqueId*);
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int);
  ncclResult_t (*CommDestroy)(ncclComm_t);
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
  ncclResult_t (*GroupStart)();
  ncclResult_t (*GroupEnd)();
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t);
  const char* (*GetErrorString)(ncclResult_t);
};
static NcclApi g_nccl;
static std::once_flag g_nccl_once;
static bool g_nccl_ok = false;
static std::string g_nccl_err;

static void load_nccl() {
  std::call_once(g_nccl_once, [] {
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) { g_nccl_err = std::string("cannot load libnccl.so.2: ") + dlerror(); return; }
#define B2_SYM(field, name) \
    g_nccl.field = reinterpret_cast<decltype(g_nccl.field)>(dlsym(h, name)); \
    if (!g_nccl.field) { g_nccl_err = std::string("libnccl lacks ") + name; return; }
    B2_SYM(GetUniqueId, "ncclGetUniqueId") B2_SYM(CommInitRank, "ncclCommInitRank") B2_SYM(CommDestroy, "ncclCommDestroy")
    B2_SYM(Send, "ncclSend") B2_SYM(Recv, "ncclRecv") B2_SYM(GroupStart, "ncclGroupStart") B2_SYM(GroupEnd, "ncclGroupEnd")
    B2_SYM(AllGather, "ncclAllGather") B2_SYM(GetErrorString, "ncclGetErrorString")
#undef B2_SYM
    g_nccl_ok = true;
  });
  if (!g_nccl_ok) throw Error(B2_ERR_UNSUPPORTED, g_nccl_err);
}
static void nccl_check(ncclResult_t r, const char* what) {
  if (r != ncclSuccess) throw Error(B2_ERR_CUDA, std::string(what) + ": " + g_nccl.GetErrorString(r));
}
#define NCCL_CHECK(x) nccl_check((x), #x)

struct Comm {
  ncclComm_t comm = nullptr;
  int rank = 0, world = 1;
};
static Comm* comm_from(b2_handle h) {
  if (!h) throw Error(B2_ERR_INVALID, "null communicator handle");
  return reinterpret_cast<Comm*>((intptr_t)h);
}

__global__ void bits_to_bytes_kernel(const uint32_t* __restrict__ bits, int64_t n, uint8_t* __restrict__ out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = bits ? (uint8_t)bit_get(bits, i) : 1;
}
__global__ void bytes_to_bits_kernel(const uint8_t* __restrict__ in, int64_t n, uint32_t* __restrict__ bits) {
  const int64_t nround = (n + 31) & ~(int64_t)31;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nround; i += (int64_t)gridDim.x * blockDim.x) {
    const bool v = i < n && in[i] != 0;
    const uint32_t b = __ballot_sync(0xffffffffu, v);
    if ((threadIdx.x & 31) == 0 && i < n) bits[i >> 5] = b;
  }
}
__global__ void lengths_kernel(const int32_t* __restrict__ offsets, int64_t n, int32_t* __restrict__ len) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) len[i] = offsets[i + 1] - offsets[i];
}

}  // namespace b2

using namespace b2;
extern "C" {

int b2_comm_unique_id(uint8_t* out128) {
  B2_TRY
  load_nccl();
  ncclUniqueId id;
  NCCL_CHECK(g_nccl.GetUniqueId(&id));
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  memcpy(out128, &id, 128);
  B2_CATCH
}

int b2_comm_init(const uint8_t* id128, int32_t rank, int32_t world, b2_handle* out_comm) {
  B2_TRY
  load_nccl();
  stream();  // binds the device
  std::unique_ptr<Comm> c(new Comm());
  c->rank = rank; c->world = world;
  ncclUniqueId id;
  memcpy(&id, id128, 128);
  NCCL_CHECK(g_nccl.CommInitRank(&c->comm, world, id, rank));
  *out_comm = to_handle(c.release());
  B2_CATCH
}

int b2_comm_close(b2_handle h) {
  B2_TRY
  Comm* c = comm_from(h);
  if (c->comm) g_nccl.CommDestroy(c->comm);
  delete c;
  B2_CATCH
}

int b2_exchange(b2_handle comm, b2_handle partitioned_table, const int32_t* offsets, b2_handle* out_table) {
  B2_TRY
  Comm* c = comm_from(comm);
  Table* t = table_from(partitioned_table);
  const int W = c->world, me = c->rank;
  const int ncols = (int)t->cols.size();
  cudaStream_t s = stream();
  B2_CHECK(offsets[0] == 0 && offsets[W] == t->rows, "offsets must cover the table");
  // per destination: rows, then chars per string column
  std::vector<int> str_cols;
  for (int i = 0; i < ncols; i++) if (t->cols[i]->dtype == B2_STRING) str_cols.push_back(i);
  const int S = 1 + (int)str_cols.size();
  std::vector<int64_t> send_sz((size_t)W * S, 0);
  std::vector<std::vector<int32_t>> str_off(str_cols.size());
  for (size_t k = 0; k < str_cols.size(); k++) {  // char offsets at the partition boundaries
    const Column* col = t->cols[str_cols[k]];
    str_off[k].resize(W + 1);
    for (int r = 0; r <= W; r++) d2h(&str_off[k][r], col->offsets.as<int32_t>() + offsets[r], 1);
  }
  if (!str_cols.empty()) sync();
  for (int r = 0; r < W; r++) {
    send_sz[(size_t)r * S] = offsets[r + 1] - offsets[r];
    for (size_t k = 0; k < str_cols.size(); k++) send_sz[(size_t)r * S + 1 + k] = str_off[k][r + 1] - str_off[k][r];
  }
  // size matrix: all[src][dst][S]
  DevBuf d_send((size_t)W * S * 8), d_all((size_t)W * W * S * 8);
  h2d(d_send.p, send_sz.data(), send_sz.size());
  NCCL_CHECK(g_nccl.AllGather(d_send.p, d_all.p, (size_t)W * S * 8, ncclInt8, c->comm, s));
  std::vector<int64_t> all((size_t)W * W * S);
  d2h(all.data(), d_all.p, all.size());
  sync();
  auto sz = [&](int src, int dst, int k) { return all[((size_t)src * W + dst) * S + k]; };
  std::vector<int64_t> recv_row_off(W + 1, 0);
  for (int src = 0; src < W; src++) recv_row_off[src + 1] = recv_row_off[src] + sz(src, me, 0);
  const int64_t out_rows = recv_row_off[W];
  if (out_rows > 0x7fffffffLL) throw Error(B2_ERR_SIZE_OVERFLOW, "exchange result exceeds 2^31-1 rows");

  ColsGuard outs;
  struct Xfer { const char* sbuf; char* rbuf; std::vector<int64_t> soff, roff; int64_t unit; };
  std::vector<Xfer> xfers;
  std::vector<DevBuf> temps;
  std::vector<std::pair<Column*, uint8_t*>> valid_fix;  // (column, received validity bytes)
  std::vector<std::pair<Column*, int32_t*>> str_fix;    // (column, received lengths)
  auto row_xfer = [&](const void* sbuf, void* rbuf, int64_t unit) {
    Xfer x; x.sbuf = (const char*)sbuf; x.rbuf = (char*)rbuf; x.unit = unit;
    for (int r = 0; r <= W; r++) { x.soff.push_back(offsets[r]); x.roff.push_back(recv_row_off[r]); }
    xfers.push_back(std::move(x));
  };
  int sidx = 0;
  for (int i = 0; i < ncols; i++) {
    const Column* ic = t->cols[i];
    // nullability must agree on every rank for the grouped send/recv to pair up: always ship validity bytes
    std::unique_ptr<Column> oc(new Column());
    oc->dtype = ic->dtype; oc->scale = ic->scale; oc->size = out_rows;
    oc->valid = DevBuf(validity_bytes(out_rows)); oc->null_count = -1;
    temps.emplace_back((size_t)std::max<int64_t>(t->rows, 1));
    uint8_t* sv = temps.back().as<uint8_t>();
    if (t->rows) { bits_to_bytes_kernel<<<grid_for(t->rows, 256), 256, 0, s>>>(ic->validity(), t->rows, sv); count_launch(); }
    temps.emplace_back((size_t)std::max<int64_t>(out_rows, 1));
    uint8_t* rv = temps.back().as<uint8_t>();
    row_xfer(sv, rv, 1);
    valid_fix.push_back({oc.get(), rv});
    if (ic->dtype == B2_STRING) {
      temps.emplace_back((size_t)std::max<int64_t>(t->rows, 1) * 4);
      int32_t* slen = temps.back().as<int32_t>();
      if (t->rows) { lengths_kernel<<<grid_for(t->rows, 256), 256, 0, s>>>(ic->offsets.as<int32_t>(), t->rows, slen); count_launch(); }
      temps.emplace_back((size_t)(out_rows + 1) * 4);
      int32_t* rlen = temps.back().as<int32_t>();
      row_xfer(slen, rlen, 4);
      int64_t chars = 0;
      Xfer x; x.sbuf = ic->data.as<char>(); x.unit = 1;
      x.soff.resize(W + 1); x.roff.resize(W + 1);
      for (int r = 0; r <= W; r++) x.soff[r] = str_off[sidx][r];
      x.roff[0] = 0;
      for (int src = 0; src < W; src++) { x.roff[src + 1] = x.roff[src] + sz(src, me, 1 + sidx); }
      chars = x.roff[W];
      if (chars > 0x7fffffffLL) throw Error(B2_ERR_SIZE_OVERFLOW, "exchange result exceeds 2^31-1 chars");
      oc->data = DevBuf((size_t)chars); oc->chars_bytes = chars;
      oc->offsets = DevBuf((size_t)(out_rows + 1) * 4);
      x.rbuf = oc->data.as<char>();
      xfers.push_back(std::move(x));
      str_fix.push_back({oc.get(), rlen});
      sidx++;
    } else {
      const int w = dtype_width(ic->dtype);
      oc->data = DevBuf((size_t)out_rows * w);
      row_xfer(ic->data.p, oc->data.p, w);
    }
    outs.v.push_back(oc.release());
  }
  // one grouped launch: every column slice to every peer (self included: NCCL copies locally)
  NCCL_CHECK(g_nccl.GroupStart());
  for (auto& x : xfers) {
    for (int r = 0; r < W; r++) {
      const int64_t sb = (x.soff[r + 1] - x.soff[r]) * x.unit, rb = (x.roff[r + 1] - x.roff[r]) * x.unit;
      if (sb) NCCL_CHECK(g_nccl.Send(x.sbuf + x.soff[r] * x.unit, (size_t)sb, ncclInt8, r, c->comm, s));
      if (rb) NCCL_CHECK(g_nccl.Recv(x.rbuf + x.roff[r] * x.unit, (size_t)rb, ncclInt8, r, c->comm, s));
    }
  }
  NCCL_CHECK(g_nccl.GroupEnd());
  count_launch();
  for (auto& vf : valid_fix) {
    if (out_rows) { bytes_to_bits_kernel<<<grid_for(out_rows, 256), 256, 0, s>>>(vf.second, out_rows, vf.first->valid.as<uint32_t>()); count_launch(); }
  }
  for (auto& sf : str_fix) exclusive_scan<int32_t, int32_t>(sf.second, sf.first->offsets.as<int32_t>(), out_rows, true);
  sync();  // temps are freed on return
  *out_table = to_handle(new_table(outs.release()));
  B2_CATCH
}

}  // extern "C"
