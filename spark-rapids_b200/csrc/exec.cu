// exec.cu — host-side operator layer above the kernels: the C++ mirror of the reference's GpuExec
// nodes for the hot path (the reference's host side is Scala; no JVM exists in this image, so the
// compiled-language mirror is C++).  Contract copied from GpuExec.internalDoExecuteColumnar():
// RDD[ColumnarBatch] (GpuExec.scala:106,190,380): every node is a pull iterator of batches; metrics
// numOutputRows / numOutputBatches / opTime (GpuExec.scala:195-199, GpuMetrics.scala:49-62).
//
//   GpuBatchSource            the child RDD / scan output handed in by the caller
//   GpuParquetScanExec        GpuParquetScan.scala:3543-3600 (PartitionReader.next -> Table.readParquet)
//   GpuFilterExec             basicPhysicalOperators.scala:1238-1291
//   GpuProjectExec            basicPhysicalOperators.scala:755-884
//   GpuHashAggregateExec      GpuAggregateExec.scala:1942-2085; first pass per batch (:730-742) then
//                             GpuMergeAggregateIterator (:896-1018): concat partials + merge-aggregate
//   GpuShuffledHashJoinExec   GpuShuffledHashJoinExec.scala:228-385 + GpuHashJoin.doJoin (:2547-2628):
//                             build side coalesced to ONE batch, hash table built once, stream probed
//                             batch by batch, payload gathered
//   GpuSortExec / GpuTopN     GpuSortExec.scala:87-165 (each-batch | full), limit.scala:234-330
//   GpuCoalesceBatches        GpuCoalesceBatches.scala:160-239 (TargetSize goal by rows)
//   GpuShuffleExchangeExec    GpuShuffleExchangeExecBase.scala:384-536 with the NCCL all-to-all
#include <chrono>
#include <deque>
#include "vm.cuh"

namespace b2 {

Table* gather_table(const Table* t, const int32_t* d_map, int64_t n, bool nullify_oob, const std::vector<int>* only_cols);
Table* concat_tables(const std::vector<const Table*>& ts);
Table* scan_aggregate(const Program* prog, bool has_pred, const Table* t, const int* key_outs, int nkeys, const b2_agg_spec* specs, int naggs);
Program* make_passthrough_program(const Table* t, const std::vector<int>& cols);
Table* filter_select(const Program* prog, const Table* t, const int32_t* keep, int nkeep);
Column* filter_row_ids(const Program* prog, const Table* t);
bool join_probe_pred(b2_handle ht, const Table* batch, int key_col, const Program* prog, Column** out_lm, Column** out_rm, int64_t* npass_out);   // join.cu
Table* filter_by_mask(const Table* t, Column* m);
Column* rows_with_passing_pair(const Column* left_map, const Column* pass, int64_t stream_rows, bool invert);

Table* slice_table(const Table* t, int64_t start, int64_t end);

// RmmRapidsRetryIterator.withRetry (RmmRapidsRetryIterator.scala:65-203): an allocation failure first spills (core.cu does
// that inside the allocator) and surfaces as B2_ERR_OOM = GpuRetryOOM; the operator then makes everything spillable leave
// the device, waits for the stream and runs the attempt again.  After two retries the error is handed to the caller, which
// for joins and aggregates splits the input batch in halves and retries each (GpuSplitAndRetryOOM semantics,
// AbstractGpuJoinIterator.scala:235-250 for the join's 2^31 output limit).
template <typename F>
static auto with_retry(F&& attempt) -> decltype(attempt()) {
  for (int n = 0;; n++) {
    try {
      return attempt();
    } catch (const Error& e) {
      if (e.code != B2_ERR_OOM || n >= 2) throw;
      note_retry();
      spill_device(INT64_MAX);
      CUDA_CHECK(cudaStreamSynchronize(stream()));
    }
  }
}
static bool splittable(const Error& e) { return e.code == B2_ERR_OOM || e.code == B2_ERR_SIZE_OVERFLOW; }

struct TableRef {  // owning reference
  Table* t = nullptr;
  TableRef() {}
  explicit TableRef(Table* t_) : t(t_) {}
  TableRef(const TableRef&) = delete;
  TableRef& operator=(const TableRef&) = delete;
  TableRef(TableRef&& o) noexcept : t(o.t) { o.t = nullptr; }
  TableRef& operator=(TableRef&& o) noexcept { if (this != &o) { reset(); t = o.t; o.t = nullptr; } return *this; }
  ~TableRef() { reset(); }
  void reset() { if (t) table_release(t); t = nullptr; }
  Table* release() { Table* r = t; t = nullptr; return r; }
};

static Table* from_handle_owned(b2_handle h) { return h ? reinterpret_cast<Table*>((intptr_t)h) : nullptr; }

struct GpuExec {
  std::atomic<int> refs{1};
  std::vector<GpuExec*> children;
  int64_t num_output_rows = 0, num_output_batches = 0, op_time_ns = 0;
  // device time of this node (CUDA events on the library stream around every next(), recorded while profiling is on):
  // total includes the children pulled from inside next(); self = total - children  (opTime of GpuMetrics.scala:49-62)
  struct Span { cudaEvent_t a, b; };
  std::vector<Span> spans, child_spans;
  virtual ~GpuExec() {
    for (auto& sp : spans) { cudaEventDestroy(sp.a); cudaEventDestroy(sp.b); }
    for (auto* c : children) c->release();
  }
  void release() { if (refs.fetch_sub(1) == 1) delete this; }
  void add_child(GpuExec* c) { c->refs.fetch_add(1); children.push_back(c); }
  virtual Table* do_next() = 0;  // nullptr when exhausted; returned table is owned by the caller
  static GpuExec*& current() { static thread_local GpuExec* cur = nullptr; return cur; }
  Table* next() {
    const bool prof = profile_enabled();
    Span sp{nullptr, nullptr};
    if (prof) { cudaEventCreate(&sp.a); cudaEventCreate(&sp.b); cudaEventRecord(sp.a, stream()); }
    GpuExec* parent = current();
    struct Restore { GpuExec* p; ~Restore() { current() = p; } } restore{parent};
    current() = this;
    auto t0 = std::chrono::steady_clock::now();
    Table* t = do_next();
    op_time_ns += std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - t0).count();
    if (prof) { cudaEventRecord(sp.b, stream()); spans.push_back(sp); if (parent) parent->child_spans.push_back(sp); }
    if (t) { num_output_rows += t->rows; num_output_batches++; }
    return t;
  }
  // run f() as if node `who` had done it (an operator fused into this one keeps its own line in the metrics)
  template <typename F>
  auto run_as(GpuExec* who, F&& f) -> decltype(f()) {
    if (!profile_enabled()) return f();
    Span sp{nullptr, nullptr};
    cudaEventCreate(&sp.a); cudaEventCreate(&sp.b); cudaEventRecord(sp.a, stream());
    struct Done { Span& sp; GpuExec* who; GpuExec* me; ~Done() { cudaEventRecord(sp.b, stream()); who->spans.push_back(sp); me->child_spans.push_back(sp); } } done{sp, who, this};
    return f();
  }
  static double span_ms(const std::vector<Span>& v) {
    double ms = 0;
    for (auto& sp : v) { float t = 0; cudaEventSynchronize(sp.b); if (cudaEventElapsedTime(&t, sp.a, sp.b) == cudaSuccess) ms += t; }
    return ms;
  }
};
static GpuExec* exec_from(b2_handle h) {
  if (!h) throw Error(B2_ERR_INVALID, "null exec handle");
  return reinterpret_cast<GpuExec*>((intptr_t)h);
}

struct GpuBatchSource : GpuExec {
  std::deque<Table*> q;
  ~GpuBatchSource() { for (auto* t : q) table_release(t); }
  Table* do_next() override {
    if (q.empty()) return nullptr;
    Table* t = q.front(); q.pop_front();
    return t;
  }
};

struct GpuParquetScanExec : GpuExec {
  struct Buf { const uint8_t* p; int64_t len; };
  std::deque<Buf> bufs;
  std::vector<std::string> names;
  Table* do_next() override {
    if (bufs.empty()) return nullptr;
    Buf b = bufs.front(); bufs.pop_front();
    std::vector<const char*> cn;
    for (auto& s : names) cn.push_back(s.c_str());
    b2_handle out = 0;
    int rc = b2_parquet_decode(b.p, b.len, cn.data(), (int)cn.size(), &out);
    if (rc != B2_OK) throw Error(rc, b2_last_error());
    return from_handle_owned(out);
  }
};

struct GpuFilterExec : GpuExec {
  b2_handle program;
  std::vector<int32_t> keep;   // non-empty: a column-pruning GpuProjectExec above the filter, fused (only these are compacted)
  Table* do_next() override {
    TableRef in(children[0]->next());
    if (!in.t) return nullptr;
    if (!keep.empty()) return filter_select(program_from(program), in.t, keep.data(), (int)keep.size());
    b2_handle out = 0;
    int rc = b2_filter(program, to_handle(in.t), &out);
    if (rc != B2_OK) throw Error(rc, b2_last_error());
    return from_handle_owned(out);
  }
};

// HostColumnarToGpu (HostColumnarToGpu.scala; GpuRowToColumnarExec.scala:937-998 for the row source): host columnar batches
// -> device batches.  The copy of batch k+1 runs on the copy stream while batch k is being consumed downstream
// (the reference keeps host buffers in flight the same way: GpuMultiFileReader.scala).
struct HostColumn { int32_t dtype, scale; int64_t rows; const void* data; const uint8_t* validity; const int32_t* offsets; };
struct GpuHostBatchSource : GpuExec {
  std::vector<std::vector<HostColumn>> batches;
  size_t next_batch = 0;
  struct InFlight { Table* t = nullptr; cudaEvent_t done = nullptr; };
  std::deque<InFlight> flight;
  cudaStream_t copy = nullptr;
  int depth = 2;
  ~GpuHostBatchSource() {
    for (auto& f : flight) { if (f.done) { cudaEventSynchronize(f.done); cudaEventDestroy(f.done); } if (f.t) table_release(f.t); }
    if (copy) cudaStreamDestroy(copy);
  }
  void issue() {
    const auto& b = batches[next_batch++];
    cudaStream_t s = stream();
    if (!copy) CUDA_CHECK(cudaStreamCreateWithFlags(&copy, cudaStreamNonBlocking));
    ColsGuard cols;
    for (const HostColumn& h : b) {
      std::unique_ptr<Column> c(new Column());
      c->dtype = h.dtype; c->scale = h.scale; c->size = h.rows;
      if (h.rows > 0x7fffffffLL) throw Error(B2_ERR_SIZE_OVERFLOW, "host batch of more than 2^31-1 rows");
      if (h.dtype == B2_STRING) {
        c->offsets = DevBuf((size_t)(h.rows + 1) * 4);
        c->chars_bytes = h.rows ? h.offsets[h.rows] : 0;
        c->data = DevBuf((size_t)c->chars_bytes);
      } else c->data = DevBuf((size_t)h.rows * dtype_width(h.dtype));
      if (h.validity) { c->valid = DevBuf(validity_bytes(h.rows)); c->null_count = -1; }
      cols.v.push_back(c.release());
    }
    // the buffers were allocated in compute-stream order: the copy stream may touch them only after that point
    cudaEvent_t ready;
    CUDA_CHECK(cudaEventCreateWithFlags(&ready, cudaEventDisableTiming));
    CUDA_CHECK(cudaEventRecord(ready, s));
    CUDA_CHECK(cudaStreamWaitEvent(copy, ready, 0));
    cudaEventDestroy(ready);
    for (size_t i = 0; i < b.size(); i++) {
      const HostColumn& h = b[i];
      Column* c = cols.v[i];
      if (h.dtype == B2_STRING) {
        if (h.rows) CUDA_CHECK(cudaMemcpyAsync(c->offsets.p, h.offsets, (size_t)(h.rows + 1) * 4, cudaMemcpyHostToDevice, copy));
        else CUDA_CHECK(cudaMemsetAsync(c->offsets.p, 0, 4, copy));
        if (c->chars_bytes) CUDA_CHECK(cudaMemcpyAsync(c->data.p, h.data, (size_t)c->chars_bytes, cudaMemcpyHostToDevice, copy));
      } else if (h.rows) {
        CUDA_CHECK(cudaMemcpyAsync(c->data.p, h.data, (size_t)h.rows * dtype_width(h.dtype), cudaMemcpyHostToDevice, copy));
      }
      if (h.validity) {
        CUDA_CHECK(cudaMemsetAsync(c->valid.p, 0, c->valid.bytes, copy));
        CUDA_CHECK(cudaMemcpyAsync(c->valid.p, h.validity, (size_t)((h.rows + 7) / 8), cudaMemcpyHostToDevice, copy));
      }
      h2d_bytes_total += (h.dtype == B2_STRING ? (int64_t)(h.rows + 1) * 4 + c->chars_bytes : h.rows * dtype_width(h.dtype)) + (h.validity ? (h.rows + 7) / 8 : 0);
    }
    InFlight f;
    f.t = new_table(cols.release());
    CUDA_CHECK(cudaEventCreateWithFlags(&f.done, cudaEventDisableTiming));
    CUDA_CHECK(cudaEventRecord(f.done, copy));
    flight.push_back(f);
  }
  int64_t h2d_bytes_total = 0;
  Table* do_next() override {
    while ((int)flight.size() < depth && next_batch < batches.size()) issue();
    if (flight.empty()) return nullptr;
    InFlight f = flight.front(); flight.pop_front();
    CUDA_CHECK(cudaStreamWaitEvent(stream(), f.done, 0));   // the consumer's stream is ordered after the copy; the host does not block
    cudaEventDestroy(f.done);
    if (next_batch < batches.size() && (int)flight.size() < depth) issue();
    return f.t;
  }
};
struct GpuProjectExec : GpuExec {
  b2_handle program;
  Table* do_next() override {
    TableRef in(children[0]->next());
    if (!in.t) return nullptr;
    b2_handle out = 0;
    int rc = b2_project(program, to_handle(in.t), &out);
    if (rc != B2_OK) throw Error(rc, b2_last_error());
    return from_handle_owned(out);
  }
};

// GpuExpandExec (GpuExpandExec.scala): every input batch is projected once per projection list and the results are stacked —
// the plan Spark uses for GROUPING SETS / ROLLUP / several COUNT(DISTINCT).  Output rows = rows x projections, projection-major.
struct GpuExpandExec : GpuExec {
  std::vector<b2_handle> projections;
  Table* do_next() override {
    TableRef in(children[0]->next());
    if (!in.t) return nullptr;
    std::vector<TableRef> parts;
    for (b2_handle p : projections) {
      b2_handle out = 0;
      int rc = b2_project(p, to_handle(in.t), &out);
      if (rc != B2_OK) throw Error(rc, b2_last_error());
      parts.emplace_back(from_handle_owned(out));
    }
    if (parts.size() == 1) return parts[0].release();
    std::vector<const Table*> ts;
    for (auto& p : parts) ts.push_back(p.t);
    return concat_tables(ts);
  }
};

// aggregate modes as in Spark: Partial/Complete run the update aggregates on raw input, Final merges
// partial buffers (SUM of sums, SUM of counts, MIN of mins, MAX of maxes)
struct GpuHashAggregateExec : GpuExec {
  Program* program = nullptr;  // pre-step projection (+ fused predicate as output 0) for update mode; null in merge mode
  bool has_pred = false, merge_mode = false, done = false;
  std::vector<int> keys;               // update: program outputs; merge: leading input columns
  std::vector<b2_agg_spec> aggs;       // update: columns = program outputs; merge: columns = input columns

  static std::vector<b2_agg_spec> merge_specs(const std::vector<b2_agg_spec>& a, int nkeys) {
    std::vector<b2_agg_spec> m = a;
    for (size_t i = 0; i < m.size(); i++) {
      m[i].column = nkeys + (int)i;
      if (m[i].kind == B2_AGG_COUNT || m[i].kind == B2_AGG_COUNT_ALL) { m[i].kind = B2_AGG_SUM; m[i].out_dtype = B2_INT64; m[i].out_scale = 0; }
    }
    return m;
  }
  Table* merge(const Table* t) {  // keys are the leading columns
    std::vector<int> cols, key_outs;
    for (int k = 0; k < (int)keys.size(); k++) { key_outs.push_back(k); cols.push_back(k); }
    std::vector<b2_agg_spec> m = merge_specs(aggs, (int)keys.size());
    for (auto& s : m) cols.push_back(s.column);
    std::unique_ptr<Program> p(make_passthrough_program(t, cols));
    return scan_aggregate(p.get(), false, t, key_outs.data(), (int)key_outs.size(), m.data(), (int)m.size());
  }
  // first-pass aggregation of one input batch; when the device cannot hold it the batch is split in halves, each half
  // yields its own partial (the merge pass combines them like any other pair of partials)
  void first_pass(const Table* in, std::vector<TableRef>& partials, int depth) {
    try {
      partials.emplace_back(with_retry([&] { return scan_aggregate(program, has_pred, in, keys.data(), (int)keys.size(), aggs.data(), (int)aggs.size()); }));
    } catch (const Error& e) {
      if (!splittable(e) || in->rows < 2 || depth >= 12) throw;
      note_split();
      const int64_t mid = in->rows / 2;
      { TableRef lo(slice_table(in, 0, mid)); first_pass(lo.t, partials, depth + 1); }
      { TableRef hi(slice_table(in, mid, in->rows)); first_pass(hi.t, partials, depth + 1); }
    }
  }
  Table* do_next() override {
    if (done) return nullptr;
    done = true;
    std::vector<TableRef> partials;
    while (true) {
      TableRef in(children[0]->next());
      if (!in.t) break;
      if (merge_mode) partials.emplace_back(in.release());   // inputs already are aggregation buffers
      else first_pass(in.t, partials, 0);
    }
    if (partials.empty()) {
      // a keyless aggregate over no batches still emits its initial-value row (GpuAggregateExec.scala:1107-1126)
      if (!keys.empty() || merge_mode) return nullptr;
      throw Error(B2_ERR_UNSUPPORTED, "keyless aggregate over zero input batches needs the input schema");
    }
    if (partials.size() == 1 && !merge_mode) return partials[0].release();
    std::vector<const Table*> ts;
    for (auto& p : partials) ts.push_back(p.t);
    TableRef cat(concat_tables(ts));
    return merge(cat.t);
  }
};

struct GpuShuffledHashJoinExec : GpuExec {  // children[0] = stream (left), children[1] = build (right)
  std::vector<int> stream_keys, build_keys;
  int kind = B2_JOIN_INNER;
  bool nulls_equal = false, built = false, full_done = false;
  // mixed join (GpuHashJoin.scala:335-530 mixed*JoinGatherMaps, :1556 ConditionalHashJoinIterator): an extra non-equi
  // condition, bound over [stream columns ++ build columns], decides which equi-matched pairs survive
  b2_handle condition = 0;
  bool pruned = false;                       // a column-pruning GpuProjectExec above the join, fused into the gathers
  std::vector<int> stream_out, build_out;    // pruned: the columns each side contributes (in this order)
  TableRef build_table;
  b2_handle ht = 0;
  ~GpuShuffledHashJoinExec() { if (ht) b2_join_hash_table_close(ht); }
  static Table* select(const Table* t, const std::vector<int>& idx) {
    std::vector<Column*> cols;
    for (int i : idx) { if (i < 0 || i >= (int)t->cols.size()) throw Error(B2_ERR_INVALID, "join key index out of range"); col_incref(t->cols[i]); cols.push_back(t->cols[i]); }
    return new_table(std::move(cols));
  }
  void build() {
    // prepareBuildBatchesForJoin: the whole build side becomes one batch
    std::vector<TableRef> parts;
    while (true) { TableRef b(children[1]->next()); if (!b.t) break; parts.emplace_back(b.release()); }
    if (parts.empty()) throw Error(B2_ERR_UNSUPPORTED, "empty build side needs the build schema");
    if (parts.size() == 1) build_table = std::move(parts[0]);
    else { std::vector<const Table*> ts; for (auto& p : parts) ts.push_back(p.t); build_table = TableRef(concat_tables(ts)); }
    TableRef bk(select(build_table.t, build_keys));
    int rc = b2_join_build(to_handle(bk.t), nulls_equal, &ht);
    if (rc != B2_OK) throw Error(rc, b2_last_error());
    built = true;
  }
  // A GpuFilterExec directly below the stream side is fused by late materialisation: the filter yields a selection vector
  // (row ids), the probe reads the keys through it and emits ORIGINAL row ids, the payload is gathered from the unfiltered
  // batch — the filtered copy of the stream side (TPC-H q3: 7.8 GB per pass over lineitem) is never written or re-read.
  GpuFilterExec* fused = nullptr;
  bool fusion_checked = false;
  int raw_col(int filter_out_col) const { return fused->keep.empty() ? filter_out_col : fused->keep[filter_out_col]; }
  Table* fused_next() {
    TableRef raw(fused->children[0]->next());
    if (!raw.t) return nullptr;
    std::vector<int> keys, left_cols;
    for (int k : stream_keys) keys.push_back(raw_col(k));
    // simple predicate + FK -> PK inner probe: the filter is evaluated INSIDE the probe kernel (no selection vector at all)
    struct Owned { Column* c = nullptr; ~Owned() { if (c) col_release(c); } Column* take() { Column* r = c; c = nullptr; return r; } };
    Owned plm, prm, sel;
    int64_t npass = 0;
    bool in_probe = false;
    if (kind == B2_JOIN_INNER && keys.size() == 1) {
      try { in_probe = join_probe_pred(ht, raw.t, keys[0], program_from(fused->program), &plm.c, &prm.c, &npass); }
      catch (const Error& e) { if (!splittable(e)) throw; in_probe = false; }   // memory pressure: the selection-vector path retries and splits
    }
    auto make_sel = [&] { if (!sel.c) sel.c = run_as(fused, [&] { return filter_row_ids(program_from(fused->program), raw.t); }); };
    if (!in_probe) make_sel();
    fused->num_output_rows += in_probe ? npass : sel.c->size; fused->num_output_batches++;
    if (pruned) for (int c : stream_out) left_cols.push_back(raw_col(c));
    else { const int n = fused->keep.empty() ? (int)raw.t->cols.size() : (int)fused->keep.size(); for (int c = 0; c < n; c++) left_cols.push_back(raw_col(c)); }
    try {
      return with_retry([&]() -> Table* {
        Column* lmc = nullptr; Column* rmc = nullptr;
        if (plm.c) { lmc = plm.take(); rmc = prm.take(); }   // the maps of the fused kernel (first attempt only)
        else {
          make_sel();
          TableRef sk(select(raw.t, keys));
          b2_handle lm = 0, rm = 0;
          int rc = b2_join_probe_sel(ht, to_handle(sk.t), to_handle(sel.c), kind, &lm, &rm);
          if (rc != B2_OK) throw Error(rc, b2_last_error());
          lmc = col_from(lm); rmc = rm ? col_from(rm) : nullptr;
        }
        ColGuard lmap(lmc);
        ColGuard rmap(rmc);
        TableRef left(gather_table(raw.t, lmap.c->data.as<int32_t>(), lmap.c->size, false, &left_cols));
        if (!rmap.c || (pruned && build_out.empty())) return left.release();
        TableRef right(gather_table(build_table.t, rmap.c->data.as<int32_t>(), rmap.c->size, kind == B2_JOIN_LEFT_OUTER, pruned ? &build_out : nullptr));
        if (pruned && stream_out.empty()) return right.release();
        std::vector<Column*> cols;
        for (auto*& c : left.t->cols) { cols.push_back(c); c = nullptr; }
        for (auto*& c : right.t->cols) { cols.push_back(c); c = nullptr; }
        left.t->cols.clear(); right.t->cols.clear();
        return new_table(std::move(cols));
      });
    } catch (const Error& e) {
      if (!splittable(e)) throw;
      // memory pressure / 2^31 limit: materialise the filter output after all and take the split-and-retry path
      std::vector<int> fcols;
      const int n = fused->keep.empty() ? (int)raw.t->cols.size() : (int)fused->keep.size();
      for (int c = 0; c < n; c++) fcols.push_back(raw_col(c));
      make_sel();
      TableRef ft(gather_table(raw.t, sel.c->data.as<int32_t>(), sel.c->size, false, &fcols));
      todo.emplace_back(std::move(ft), 0);
      return drain_todo();
    }
  }
  Table* do_next() override {
    if (!todo.empty()) return drain_todo();
    if (!built) build();
    if (!fusion_checked) {
      fusion_checked = true;
      if (kind != B2_JOIN_FULL_OUTER && !condition && !getenv("B2_NO_FILTER_FUSION")) fused = dynamic_cast<GpuFilterExec*>(children[0]);
    }
    if (fused) return fused_next();
    TableRef s;
    if (kind == B2_JOIN_FULL_OUTER) {
      // the unmatched build rows can only be emitted once every stream row has been seen: the stream side is
      // coalesced into one batch (the reference tracks matched build rows across batches instead)
      if (full_done) return nullptr;
      full_done = true;
      std::vector<TableRef> parts;
      while (true) { TableRef b(children[0]->next()); if (!b.t) break; parts.emplace_back(b.release()); }
      if (parts.empty()) throw Error(B2_ERR_UNSUPPORTED, "empty stream side of a full outer join needs the stream schema");
      if (parts.size() == 1) s = std::move(parts[0]);
      else { std::vector<const Table*> ts; for (auto& q : parts) ts.push_back(q.t); s = TableRef(concat_tables(ts)); }
    } else {
      s = TableRef(children[0]->next());
    }
    if (!s.t) return nullptr;
    todo.emplace_back(std::move(s), 0);
    return drain_todo();
  }
  // stream slices waiting to be joined (a batch that had to be split leaves its halves here; one output per next())
  std::deque<std::pair<TableRef, int>> todo;
  // one stream batch -> output batch(es).  When the gather maps would pass 2^31-1 rows (B2_ERR_SIZE_OVERFLOW) or the device
  // cannot hold the output (B2_ERR_OOM after retries) the stream batch is halved and each half joined on its own — lazily,
  // so that at most one output batch exists at a time (GpuSplitAndRetryOOM, AbstractGpuJoinIterator.scala:235-250)
  Table* drain_todo() {
    while (!todo.empty()) {
      TableRef st = std::move(todo.front().first);
      const int depth = todo.front().second;
      todo.pop_front();
      try {
        return with_retry([&] { return join_batch(st.t); });
      } catch (const Error& e) {
        if (!splittable(e) || st.t->rows < 2 || depth >= 16 || kind == B2_JOIN_FULL_OUTER) throw;
        note_split();
        const int64_t mid = st.t->rows / 2;
        TableRef lo, hi;
        // the halves replace the batch: the batch itself is released before they are made, so the split needs no extra room
        {
          std::vector<int> all;
          for (int c = 0; c < (int)st.t->cols.size(); c++) all.push_back(c);
          lo = TableRef(slice_table(st.t, 0, mid));
          hi = TableRef(slice_table(st.t, mid, st.t->rows));
        }
        st.reset();
        todo.emplace_front(std::move(hi), depth + 1);
        todo.emplace_front(std::move(lo), depth + 1);
      }
    }
    return nullptr;
  }
  static Table* concat_cols(Table* a, Table* b) {   // steals the columns of both
    std::vector<Column*> cols;
    for (auto*& c : a->cols) { cols.push_back(c); c = nullptr; }
    for (auto*& c : b->cols) { cols.push_back(c); c = nullptr; }
    a->cols.clear(); b->cols.clear();
    return new_table(std::move(cols));
  }
  Table* prune_pairs(const Table* pairs, int nstream) {   // [stream ++ build] -> the output columns of the node
    std::vector<Column*> cols;
    auto take = [&](int i) { col_incref(pairs->cols[i]); cols.push_back(pairs->cols[i]); };
    if (pruned) { for (int c : stream_out) take(c); for (int c : build_out) take(nstream + c); }
    else for (int i = 0; i < (int)pairs->cols.size(); i++) take(i);
    return new_table(std::move(cols));
  }
  // equi-join pairs -> condition over the pair rows -> per join type
  Table* join_batch_conditional(const Table* st) {
    if (kind == B2_JOIN_FULL_OUTER) throw Error(B2_ERR_UNSUPPORTED, "full outer join with a non-equi condition");
    TableRef sk(select(st, stream_keys));
    b2_handle lm = 0, rm = 0;
    int rc = b2_join_probe(ht, to_handle(sk.t), B2_JOIN_INNER, &lm, &rm);
    if (rc != B2_OK) throw Error(rc, b2_last_error());
    ColGuard lmap(col_from(lm)), rmap(col_from(rm));
    const int nstream = (int)st->cols.size();
    TableRef left(gather_table(st, lmap.c->data.as<int32_t>(), lmap.c->size, false, nullptr));
    TableRef right(gather_table(build_table.t, rmap.c->data.as<int32_t>(), rmap.c->size, false, nullptr));
    TableRef pairs(concat_cols(left.t, right.t));
    b2_handle ph = 0;
    rc = b2_project(condition, to_handle(pairs.t), &ph);
    if (rc != B2_OK) throw Error(rc, b2_last_error());
    TableRef pt(from_handle_owned(ph));
    Column* pass = pt.t->cols[0];
    if (pass->dtype != B2_BOOL8) throw Error(B2_ERR_INVALID, "join condition must be BOOL8");
    auto stream_side = [&](const Table* t) {   // semi / anti output: the node's stream columns
      std::vector<Column*> cols;
      if (pruned) for (int c : stream_out) { col_incref(t->cols[c]); cols.push_back(t->cols[c]); }
      else for (auto* c : t->cols) { col_incref(c); cols.push_back(c); }
      return new_table(std::move(cols));
    };
    if (kind == B2_JOIN_INNER) {
      TableRef out(prune_pairs(pairs.t, nstream));
      return filter_by_mask(out.t, pass);
    }
    if (kind == B2_JOIN_LEFT_SEMI || kind == B2_JOIN_LEFT_ANTI) {
      ColGuard flags(rows_with_passing_pair(lmap.c, pass, st->rows, kind == B2_JOIN_LEFT_ANTI));
      TableRef side(stream_side(st));
      return filter_by_mask(side.t, flags.c);
    }
    // LEFT OUTER: the passing pairs, then every stream row without one, NULL on the build side
    TableRef pruned_pairs(prune_pairs(pairs.t, nstream));
    TableRef matched(filter_by_mask(pruned_pairs.t, pass));
    ColGuard lonely(rows_with_passing_pair(lmap.c, pass, st->rows, true));
    TableRef lonely_stream(filter_by_mask(st, lonely.c));
    ColGuard oob(new_column(B2_INT32, 0, lonely_stream.t->rows, false));
    if (lonely_stream.t->rows) CUDA_CHECK(cudaMemsetAsync(oob.c->data.p, 0x80, (size_t)lonely_stream.t->rows * 4, stream()));   // 0x80808080 < 0: out of bounds -> NULL row
    TableRef nulls(gather_table(build_table.t, oob.c->data.as<int32_t>(), lonely_stream.t->rows, true, nullptr));
    TableRef lonely_pairs(concat_cols(lonely_stream.t, nulls.t));
    TableRef lonely_out(prune_pairs(lonely_pairs.t, nstream));
    // nullability must agree for the concatenation: concat_tables merges validity
    std::vector<const Table*> ts{matched.t, lonely_out.t};
    return concat_tables(ts);
  }
  Table* join_batch(const Table* st) {
    if (condition) return join_batch_conditional(st);
    TableRef sk(select(st, stream_keys));
    b2_handle lm = 0, rm = 0;
    int rc = b2_join_probe(ht, to_handle(sk.t), kind, &lm, &rm);
    if (rc != B2_OK) throw Error(rc, b2_last_error());
    ColGuard lmap(col_from(lm));
    ColGuard rmap(rm ? col_from(rm) : nullptr);
    TableRef left(gather_table(st, lmap.c->data.as<int32_t>(), lmap.c->size, kind == B2_JOIN_FULL_OUTER, pruned ? &stream_out : nullptr));
    if (!rmap.c || (pruned && build_out.empty())) return left.release();  // semi / anti (or nothing wanted from the build side): stream columns only
    TableRef right(gather_table(build_table.t, rmap.c->data.as<int32_t>(), rmap.c->size, kind == B2_JOIN_LEFT_OUTER || kind == B2_JOIN_FULL_OUTER,
                                pruned ? &build_out : nullptr));
    if (pruned && stream_out.empty()) return right.release();
    std::vector<Column*> cols;  // output = left columns ++ right columns (GpuHashJoin.scala:2451-2470)
    for (auto*& c : left.t->cols) { cols.push_back(c); c = nullptr; }
    for (auto*& c : right.t->cols) { cols.push_back(c); c = nullptr; }
    left.t->cols.clear(); right.t->cols.clear();
    return new_table(std::move(cols));
  }
};

DevBuf sort_order(const Table* t, const b2_order_by_arg* keys, int nkeys);
Table* top_n_table(const Table* t, const b2_order_by_arg* keys, int nkeys, int64_t limit);
struct GpuSortExec : GpuExec {
  std::vector<b2_order_by_arg> order;
  bool global = true;   // false: sort each batch (SortEachBatch); true: full sort of the partition
  int64_t limit = -1;   // >= 0: GpuTopN
  bool done = false;
  Table* sorted(const Table* t, int64_t n) {
    if (n >= 0) return top_n_table(t, order.data(), (int)order.size(), n);   // GpuTopN: radix select, then a sort of the candidates
    DevBuf perm = sort_order(t, order.data(), (int)order.size());
    return gather_table(t, perm.as<int32_t>(), n < 0 ? t->rows : std::min<int64_t>(n, t->rows), false, nullptr);
  }
  Table* do_next() override {
    if (!global && limit < 0) {
      TableRef in(children[0]->next());
      return in.t ? sorted(in.t, -1) : nullptr;
    }
    if (done) return nullptr;
    done = true;
    TableRef pending;
    while (true) {
      TableRef in(children[0]->next());
      if (!in.t) break;
      if (limit >= 0) {
        // GpuTopN: sort the batch, keep N, fold into the running top-N
        TableRef top(sorted(in.t, limit));
        if (!pending.t) pending = std::move(top);
        else { std::vector<const Table*> ts{pending.t, top.t}; TableRef cat(concat_tables(ts)); pending = TableRef(sorted(cat.t, limit)); }
      } else {
        if (!pending.t) pending = std::move(in);
        else { std::vector<const Table*> ts{pending.t, in.t}; pending = TableRef(concat_tables(ts)); }
      }
    }
    if (!pending.t) return nullptr;
    return limit >= 0 ? pending.release() : sorted(pending.t, -1);
  }
};

struct GpuCoalesceBatches : GpuExec {
  int64_t target_rows = 1 << 30;
  TableRef carry;
  Table* do_next() override {
    std::vector<TableRef> acc;
    int64_t rows = 0;
    if (carry.t) { rows += carry.t->rows; acc.emplace_back(carry.release()); }
    while (rows < target_rows) {
      TableRef in(children[0]->next());
      if (!in.t) break;
      if (rows > 0 && rows + in.t->rows > target_rows) { carry = std::move(in); break; }
      rows += in.t->rows;
      acc.emplace_back(in.release());
    }
    if (acc.empty()) return nullptr;
    if (acc.size() == 1) return acc[0].release();
    std::vector<const Table*> ts;
    for (auto& a : acc) ts.push_back(a.t);
    return concat_tables(ts);
  }
};

extern "C" int b2_exchange_hash(b2_handle comm, b2_handle table, const int32_t* key_cols, int32_t nkeys, int32_t seed, b2_handle* out_table, int32_t* any_data);
extern "C" int b2_exchange_ex(b2_handle comm, b2_handle partitioned_table, const int32_t* offsets, b2_handle* out_table, int32_t* any_data);
extern "C" int b2_comm_fused_ready(b2_handle comm, int32_t* ok);
extern "C" int b2_comm_allmax(b2_handle comm, int32_t value, int32_t* out);
extern "C" int b2_exchange_hash_sel(b2_handle comm, b2_handle table, b2_handle selection, const int32_t* out_cols, int32_t nout, const int32_t* key_cols,
                                    int32_t nkeys, int32_t seed, b2_handle* out_table, int32_t* any_data);
extern "C" int b2_broadcast_table(b2_handle comm, b2_handle table, int32_t root, b2_handle* out_table);

// GpuShuffleExchangeExec (GpuShuffleExchangeExecBase.scala:384-536).  Every call of the exchange is a collective, so the
// node follows a termination protocol instead of stopping when ITS child is exhausted: a rank without a batch keeps
// taking part (sending nothing) until no rank has data left — ranks with different batch counts, or with none at all,
// cannot strand their peers (the reference's pull-based shuffle has no such constraint).
struct GpuShuffleExchangeExec : GpuExec {
  std::vector<int32_t> key_cols;   // empty = SinglePartition (everything to rank 0)
  b2_handle comm = 0;
  int world = 1;
  bool finished = false, probed = false, fused = false;
  Table* do_next() override {
    if (finished) return nullptr;
    if (world == 1 || !comm) {
      TableRef in(children[0]->next());
      if (!in.t) { finished = true; return nullptr; }
      if (key_cols.empty()) return in.release();
      std::vector<int32_t> offs(world + 1, 0);
      b2_handle out = 0;
      int rc = b2_hash_partition(to_handle(in.t), key_cols.data(), (int)key_cols.size(), 42, world, &out, offs.data());
      if (rc != B2_OK) throw Error(rc, b2_last_error());
      return from_handle_owned(out);
    }
    // a GpuFilterExec (with its column pruning) directly below is fused into the scatter: see GpuShuffledHashJoinExec::fused
    GpuFilterExec* ff = getenv("B2_NO_FILTER_FUSION") ? nullptr : dynamic_cast<GpuFilterExec*>(children[0]);
    GpuExec* src = ff ? ff->children[0] : children[0];
    // one batch of look-ahead: the call that carries a rank's LAST batch says so (b2_comm_set_more), and when no rank has
    // more the exchange is over without an extra, empty round (one header all-gather + D2H + sync per exchange node)
    if (!primed) { primed = true; ahead = TableRef(src->next()); }
    TableRef in = std::move(ahead);
    if (in.t) ahead = TableRef(src->next());
    {
      int rc = b2_comm_set_more(comm, ahead.t ? 1 : 0);
      if (rc != B2_OK) throw Error(rc, b2_last_error());
    }
    std::vector<int32_t> out_cols;     // the columns that travel (indices into the batch pulled from src)
    if (ff && in.t) {
      if (ff->keep.empty()) for (int c = 0; c < (int)in.t->cols.size(); c++) out_cols.push_back(c);
      else out_cols = ff->keep;
    }
    if (!probed) {   // collective, once per node: can every rank store into every peer's arena, and is the schema fixed width?
      int32_t ok = 0;
      int rc = b2_comm_fused_ready(comm, &ok);
      if (rc != B2_OK) throw Error(rc, b2_last_error());
      // a rank without a batch does not know the schema, so the ranks agree on "some batch carries STRING columns"
      int32_t mine = 0, any_strings = 0;
      if (in.t && !ff) for (auto* c : in.t->cols) if (c->dtype == B2_STRING) mine = 1;
      if (in.t && ff) for (int c : out_cols) if (in.t->cols[c]->dtype == B2_STRING) mine = 1;
      rc = b2_comm_allmax(comm, mine, &any_strings);
      if (rc != B2_OK) throw Error(rc, b2_last_error());
      has_strings = any_strings != 0;
      fused = ok != 0; probed = true;
    }
    b2_handle out = 0;
    int32_t any = 0;
    if (fused && !has_strings && ff) {
      ColGuard sel(in.t ? run_as(ff, [&] { return filter_row_ids(program_from(ff->program), in.t); }) : nullptr);
      if (in.t) { ff->num_output_rows += sel.c->size; ff->num_output_batches++; }
      std::vector<int32_t> keys;   // key_cols index the filter's output columns
      for (int32_t k : key_cols) keys.push_back(ff->keep.empty() ? k : ff->keep[k]);
      int rc = b2_exchange_hash_sel(comm, in.t ? to_handle(in.t) : 0, sel.c ? to_handle(sel.c) : 0, out_cols.data(), (int)out_cols.size(), keys.data(),
                                    (int)keys.size(), 42, &out, &any);
      if (rc != B2_OK) throw Error(rc, b2_last_error());
    } else if (fused && !has_strings) {
      int rc = b2_exchange_hash(comm, in.t ? to_handle(in.t) : 0, key_cols.data(), (int)key_cols.size(), 42, &out, &any);
      if (rc != B2_OK) throw Error(rc, b2_last_error());
    } else {
      if (ff && in.t) {   // NCCL path: the filter runs unfused
        TableRef f(ff->keep.empty() ? nullptr : filter_select(program_from(ff->program), in.t, ff->keep.data(), (int)ff->keep.size()));
        if (!f.t) { b2_handle fh = 0; int rc = b2_filter(ff->program, to_handle(in.t), &fh); if (rc != B2_OK) throw Error(rc, b2_last_error()); f = TableRef(from_handle_owned(fh)); }
        ff->num_output_rows += f.t->rows; ff->num_output_batches++;
        in = std::move(f);
      }
      std::vector<int32_t> offs(world + 1, 0);
      TableRef part;
      if (in.t) {
        if (key_cols.empty()) {
          in.t->refs.fetch_add(1);
          part = TableRef(in.t);
          for (int r = 1; r <= world; r++) offs[r] = (int32_t)in.t->rows;
        } else {
          b2_handle ph = 0;
          int rc = b2_hash_partition(to_handle(in.t), key_cols.data(), (int)key_cols.size(), 42, world, &ph, offs.data());
          if (rc != B2_OK) throw Error(rc, b2_last_error());
          part = TableRef(from_handle_owned(ph));
        }
      }
      int rc = b2_exchange_ex(comm, part.t ? to_handle(part.t) : 0, part.t ? offs.data() : nullptr, &out, &any);
      if (rc != B2_OK) throw Error(rc, b2_last_error());
    }
    int32_t any_more = 1;
    {
      int rc = b2_comm_any_more(comm, &any_more);
      if (rc != B2_OK) throw Error(rc, b2_last_error());
    }
    if (!any_more) finished = true;            // every rank sent its last batch (or had none): no further round
    if (!any) { finished = true; if (out) table_release(from_handle_owned(out)); return nullptr; }
    return from_handle_owned(out);
  }
  bool primed = false;
  TableRef ahead;             // the batch the NEXT call will carry
  bool has_strings = false;   // the path (fused / NCCL) is a per-node property every rank agrees on at the first call
};

// GpuBroadcastExchangeExec (GpuBroadcastExchangeExec.scala:1-664): the child relation is collected once and every rank gets
// the WHOLE relation.  SPMD form: each rank coalesces its slice and broadcasts it (ncclBroadcast per rank), the result is
// the concatenation in rank order.  One batch out.
struct GpuBroadcastExchangeExec : GpuExec {
  b2_handle comm = 0;
  int world = 1, rank = 0;
  bool done = false;
  std::vector<int> schema_dtype, schema_scale;   // needed when this rank's slice is empty
  Table* do_next() override {
    if (done) return nullptr;
    done = true;
    std::vector<TableRef> parts;
    while (true) { TableRef b(children[0]->next()); if (!b.t) break; parts.emplace_back(b.release()); }
    TableRef mine;
    if (parts.size() == 1) mine = std::move(parts[0]);
    else if (parts.size() > 1) { std::vector<const Table*> ts; for (auto& p : parts) ts.push_back(p.t); mine = TableRef(concat_tables(ts)); }
    if (world == 1 || !comm) return mine.release();
    if (!mine.t) throw Error(B2_ERR_UNSUPPORTED, "broadcast of a slice with no batches needs the schema (push an empty batch)");
    std::vector<TableRef> got;
    for (int r = 0; r < world; r++) {
      b2_handle out = 0;
      int rc = b2_broadcast_table(comm, r == rank ? to_handle(mine.t) : 0, r, &out);
      if (rc != B2_OK) throw Error(rc, b2_last_error());
      got.emplace_back(from_handle_owned(out));
    }
    std::vector<const Table*> ts;
    for (auto& g : got) ts.push_back(g.t);
    return concat_tables(ts);
  }
};

}  // namespace b2

using namespace b2;
extern "C" {

int b2_exec_source(b2_handle* out) {
  B2_TRY
  *out = to_handle(new GpuBatchSource());
  B2_CATCH
}
int b2_exec_source_push(b2_handle src, b2_handle table) {
  B2_TRY
  auto* s = dynamic_cast<GpuBatchSource*>(exec_from(src));
  B2_CHECK(s, "not a batch source");
  Table* t = table_from(table);
  t->refs.fetch_add(1);
  s->q.push_back(t);
  B2_CATCH
}
int b2_exec_parquet_scan(const char* const* column_names, int32_t ncols, b2_handle* out) {
  B2_TRY
  auto* e = new GpuParquetScanExec();
  for (int i = 0; i < ncols; i++) e->names.push_back(column_names[i]);
  *out = to_handle(e);
  B2_CATCH
}
int b2_exec_parquet_scan_add(b2_handle scan, const uint8_t* host_buf, int64_t len) {
  B2_TRY
  auto* s = dynamic_cast<GpuParquetScanExec*>(exec_from(scan));
  B2_CHECK(s, "not a parquet scan");
  s->bufs.push_back({host_buf, len});
  B2_CATCH
}
int b2_exec_filter(b2_handle child, b2_handle predicate_program, b2_handle* out) {
  B2_TRY
  auto* e = new GpuFilterExec();
  e->add_child(exec_from(child)); e->program = predicate_program;
  *out = to_handle(e);
  B2_CATCH
}
int b2_exec_filter_select(b2_handle child, b2_handle predicate_program, const int32_t* keep_cols, int32_t nkeep, b2_handle* out) {
  B2_TRY
  B2_CHECK(nkeep >= 1, "filter: at least one output column");
  auto* e = new GpuFilterExec();
  e->add_child(exec_from(child)); e->program = predicate_program;
  e->keep.assign(keep_cols, keep_cols + nkeep);
  *out = to_handle(e);
  B2_CATCH
}
int b2_exec_host_source(b2_handle* out) {
  B2_TRY
  *out = to_handle(new GpuHostBatchSource());
  B2_CATCH
}
int b2_exec_host_source_push(b2_handle source, const b2_host_column* cols, int32_t ncols) {
  B2_TRY
  auto* s = dynamic_cast<GpuHostBatchSource*>(exec_from(source));
  B2_CHECK(s, "not a host batch source");
  B2_CHECK(ncols >= 1, "a batch needs columns");
  std::vector<HostColumn> b;
  for (int i = 0; i < ncols; i++) {
    B2_CHECK(cols[i].rows == cols[0].rows, "host batch columns differ in length");
    b.push_back({cols[i].dtype, cols[i].scale, cols[i].rows, cols[i].data, cols[i].validity_bits, cols[i].offsets});
  }
  s->batches.push_back(std::move(b));
  B2_CATCH
}
int b2_exec_expand(b2_handle child, const b2_handle* projection_programs, int32_t nprojections, b2_handle* out) {
  B2_TRY
  B2_CHECK(nprojections >= 1, "expand needs at least one projection");
  auto* e = new GpuExpandExec();
  e->add_child(exec_from(child));
  e->projections.assign(projection_programs, projection_programs + nprojections);
  *out = to_handle(e);
  B2_CATCH
}
int b2_exec_project(b2_handle child, b2_handle program, b2_handle* out) {
  B2_TRY
  auto* e = new GpuProjectExec();
  e->add_child(exec_from(child)); e->program = program;
  *out = to_handle(e);
  B2_CATCH
}
int b2_exec_hash_aggregate(b2_handle child, b2_handle program, int32_t has_predicate, int32_t merge_mode, const int32_t* keys, int32_t nkeys,
                           const b2_agg_spec* aggs, int32_t naggs, b2_handle* out) {
  B2_TRY
  auto* e = new GpuHashAggregateExec();
  e->add_child(exec_from(child));
  e->program = merge_mode ? nullptr : program_from(program);
  e->has_pred = has_predicate != 0; e->merge_mode = merge_mode != 0;
  e->keys.assign(keys, keys + nkeys); e->aggs.assign(aggs, aggs + naggs);
  *out = to_handle(e);
  B2_CATCH
}
int b2_exec_shuffled_hash_join(b2_handle stream_child, b2_handle build_child, const int32_t* stream_keys, const int32_t* build_keys, int32_t nkeys,
                               int32_t kind, int32_t nulls_equal, b2_handle* out) {
  B2_TRY
  auto* e = new GpuShuffledHashJoinExec();
  e->add_child(exec_from(stream_child)); e->add_child(exec_from(build_child));
  e->stream_keys.assign(stream_keys, stream_keys + nkeys); e->build_keys.assign(build_keys, build_keys + nkeys);
  e->kind = kind; e->nulls_equal = nulls_equal != 0;
  *out = to_handle(e);
  B2_CATCH
}
int b2_exec_shuffled_hash_join_select(b2_handle stream_child, b2_handle build_child, const int32_t* stream_keys, const int32_t* build_keys, int32_t nkeys,
                                      int32_t kind, int32_t nulls_equal, const int32_t* stream_out, int32_t nstream_out, const int32_t* build_out,
                                      int32_t nbuild_out, b2_handle* out) {
  B2_TRY
  B2_CHECK(nstream_out + nbuild_out >= 1, "join: at least one output column");
  auto* e = new GpuShuffledHashJoinExec();
  e->add_child(exec_from(stream_child)); e->add_child(exec_from(build_child));
  e->stream_keys.assign(stream_keys, stream_keys + nkeys); e->build_keys.assign(build_keys, build_keys + nkeys);
  e->kind = kind; e->nulls_equal = nulls_equal != 0;
  e->pruned = true;
  e->stream_out.assign(stream_out, stream_out + nstream_out); e->build_out.assign(build_out, build_out + nbuild_out);
  *out = to_handle(e);
  B2_CATCH
}
int b2_exec_join_set_condition(b2_handle join, b2_handle condition_program) {
  B2_TRY
  auto* j = dynamic_cast<GpuShuffledHashJoinExec*>(exec_from(join));
  B2_CHECK(j, "not a hash join node");
  j->condition = condition_program;
  B2_CATCH
}
int b2_exec_broadcast_exchange(b2_handle child, b2_handle comm, int32_t rank, int32_t world, b2_handle* out) {
  B2_TRY
  auto* e = new GpuBroadcastExchangeExec();
  e->add_child(exec_from(child));
  e->comm = comm; e->rank = rank; e->world = world;
  *out = to_handle(e);
  B2_CATCH
}
int b2_exec_sort(b2_handle child, const b2_order_by_arg* order, int32_t norder, int32_t global, int64_t limit, b2_handle* out) {
  B2_TRY
  auto* e = new GpuSortExec();
  e->add_child(exec_from(child));
  e->order.assign(order, order + norder); e->global = global != 0; e->limit = limit;
  *out = to_handle(e);
  B2_CATCH
}
int b2_exec_coalesce(b2_handle child, int64_t target_rows, b2_handle* out) {
  B2_TRY
  auto* e = new GpuCoalesceBatches();
  e->add_child(exec_from(child)); e->target_rows = target_rows;
  *out = to_handle(e);
  B2_CATCH
}
int b2_exec_shuffle_exchange(b2_handle child, const int32_t* key_cols, int32_t nkeys, b2_handle comm, int32_t world, b2_handle* out) {
  B2_TRY
  auto* e = new GpuShuffleExchangeExec();
  e->add_child(exec_from(child));
  e->key_cols.assign(key_cols, key_cols + nkeys); e->comm = comm; e->world = world;
  *out = to_handle(e);
  B2_CATCH
}
int b2_exec_next(b2_handle exec, b2_handle* out_table) {
  B2_TRY
  semaphore_acquire_if_necessary();   // GpuSemaphore.acquireIfNecessary: the task enters the GPU at its first batch
  *out_table = to_handle(exec_from(exec)->next());
  B2_CATCH
}
int b2_exec_metrics(b2_handle exec, int64_t* out3) {
  B2_TRY
  GpuExec* e = exec_from(exec);
  out3[0] = e->num_output_rows; out3[1] = e->num_output_batches; out3[2] = e->op_time_ns;
  B2_CATCH
}
int b2_exec_device_time(b2_handle exec, double* out2) {
  B2_TRY
  GpuExec* e = exec_from(exec);
  const double total = GpuExec::span_ms(e->spans), kids = GpuExec::span_ms(e->child_spans);
  out2[0] = total - kids; out2[1] = total;
  B2_CATCH
}
int b2_exec_close(b2_handle exec) {
  B2_TRY
  exec_from(exec)->release();
  B2_CATCH
}

}  // extern "C"
