// join.cu — a6: hash join build / probe producing gather maps.
// Reference: JoinImpl.innerHashJoin{BuildLeft,BuildRight} -> JoinPrimitives.hashInnerJoin
// (GpuHashJoin.scala:309-409), leftJoin/leftSemi/leftAnti gather maps (:466-600),
// JoinPrimitives.makeLeftOuter/makeSemi/makeAnti (:256-302), null-key rules (:602-640),
// HashJoinIterator (:1374-1553), build-side null filtering (GpuShuffledHashJoinExec.scala:207-212).
//
// The reference API takes both key tables on every call, so the hash table is rebuilt for every
// stream batch.  Here the build side is hashed ONCE into a persistent open-addressing multimap
// (8-byte slots: 32-bit hash tag + 32-bit build row) that any number of probe calls reuse.
// Probe is count -> scan -> write so the gather maps are exactly sized and ordered by stream row.
#include <type_traits>
#include "prim.cuh"
#include "rowops.cuh"
#include "simplefilter.cuh"

namespace b2 {

constexpr uint64_t JSLOT_EMPTY = 0xffffffffffffffffull;

struct JoinTable {
  Table* keys;        // one reference on the build key table
  DevBuf slots;       // uint64 [cap]
  int64_t cap;
  int64_t build_rows;
  bool nulls_equal;
  bool fast = false;      // keys fixed width, <= 8 bytes together, no NULL can match: 16-byte entries {tag|row, packed key}
  bool distinct = false;  // no two build rows share a key (FK -> PK joins): probe is single pass
  // Blocked Bloom filter over the build keys (two bits of one 64-bit word per key, <= 64 MB so that it stays L2 resident
  // while the table itself — 16 B per slot at load <= 0.5 — lives in HBM): a probe row whose key is absent from the build
  // side (most rows of a selective join: TPC-H q3 matches 10-20 %) is rejected by ONE L2 hit instead of a random HBM
  // access into the table.  Same idea as the runtime Bloom filter Spark injects in front of such joins
  // (GpuBloomFilterMightContain in the reference), applied inside the probe.
  DevBuf bloom;
  uint32_t bloom_mask = 0;  // 64-bit words - 1; 0 = no filter
  std::vector<int> key_idx;
  ~JoinTable() { if (keys) table_release(keys); }
};

__device__ __forceinline__ bool any_null_key(const KeyCols& k, int64_t r) {
  for (int i = 0; i < k.n; i++) if (!row_valid(k.c[i].valid, r)) return true;
  return false;
}

__device__ __forceinline__ uint64_t pack_join_key(const KeyCols& ks, int64_t r) {
  uint64_t bits = 0; int shift = 0;
  for (int i = 0; i < ks.n; i++) { bits |= key_bits(ks.c[i], r) << shift; shift += 8 * ks.c[i].width; }
  return bits;
}
__device__ __forceinline__ uint32_t hash_packed(uint64_t kb) { const uint64_t h = mix64(kb ^ 0x9e3779b97f4a7c15ull); return (uint32_t)(h ^ (h >> 32)); }

// bloom_mask: bits 0..27 = words - 1, bits 28..29 = extra bits set per key beyond the first two (k = 2..4)
__device__ __forceinline__ void bloom_of(uint32_t h, uint32_t bloom_mask, uint32_t& word, unsigned long long& bits) {
  const uint64_t p = (uint64_t)h * 0x9E3779B97F4A7C15ull;
  word = (uint32_t)(p >> 40) & (bloom_mask & 0x0fffffffu);
  bits = (1ull << ((p >> 8) & 63)) | (1ull << ((p >> 14) & 63));
  const uint32_t extra = bloom_mask >> 28;
  if (extra >= 1) bits |= 1ull << ((p >> 20) & 63);
  if (extra >= 2) bits |= 1ull << ((p >> 26) & 63);
}
__device__ __forceinline__ bool bloom_may_contain(const unsigned long long* __restrict__ bloom, uint32_t bloom_mask, uint32_t h) {
  if (!bloom) return true;
  uint32_t w; unsigned long long bits;
  bloom_of(h, bloom_mask, w, bits);
  return (__ldg(&bloom[w]) & bits) == bits;
}

__global__ void join_build_kernel(const __grid_constant__ KeyCols keys, int64_t n, uint64_t* __restrict__ slots,
                                  uint32_t mask, bool nulls_equal, bool fast, int32_t* __restrict__ has_dups,
                                  unsigned long long* __restrict__ bloom, uint32_t bloom_mask) {
  const int sh = fast ? 1 : 0;  // entry stride: 2 words when the packed key rides along
  for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < n; r += (int64_t)gridDim.x * blockDim.x) {
    if (!nulls_equal && any_null_key(keys, r)) continue;  // a NULL key can never match: keep it out of the table
    uint64_t kb = 0;
    uint32_t h;
    if (fast) { kb = pack_join_key(keys, r); h = hash_packed(kb); } else h = row_hash(keys, r);
    const uint64_t entry = ((uint64_t)h << 32) | (uint32_t)r;
    if (bloom) { uint32_t w; unsigned long long bits; bloom_of(h, bloom_mask, w, bits); atomicOr(&bloom[w], bits); }
    uint32_t idx = h & mask;
    while (true) {
      unsigned long long old = atomicCAS(reinterpret_cast<unsigned long long*>(&slots[(size_t)idx << sh]), (unsigned long long)JSLOT_EMPTY, (unsigned long long)entry);
      if (old == JSLOT_EMPTY) { if (fast) slots[((size_t)idx << 1) + 1] = kb; break; }
      // occupied: a slot with my tag may hold my key -> the build side is not distinct
      if ((uint32_t)(old >> 32) == h && *has_dups == 0 && rows_equal(keys, r, keys, (int32_t)(uint32_t)old, nulls_equal)) atomicExch(has_dups, 1);
      idx = (idx + 1) & mask;
    }
  }
}

// one probe step: entry word (and, for packed keys, the key from the same 16-byte line)
__device__ __forceinline__ uint64_t join_entry(const uint64_t* __restrict__ slots, uint32_t idx, bool fast, uint64_t& key) {
  if (fast) { const ulonglong2 e = *reinterpret_cast<const ulonglong2*>(&slots[(size_t)idx << 1]); key = e.y; return e.x; }
  key = 0;
  return slots[idx];
}

// single pass for a distinct build side: at most one match per stream row.
//   INNER: pairs appended with ONE atomic per warp and PI x 32 rows (join output order is unspecified: docs/compatibility.md:18-25)
//   LEFT OUTER: row r -> (r, match or INT32_MIN), no compaction at all
// Each warp owns chunks of PI x 32 consecutive rows and works on them in phases, so that every lane has PI independent
// loads in flight at each step (selection vector -> keys -> Bloom words -> table slots) instead of one dependent chain, and
// the output range of a whole chunk is reserved with a single atomic (the per-32-rows atomic on one address was the
// serialisation point of the kernel: ~10 M same-address atomics per TPC-H q3 step).
constexpr int PI = 8;
__global__ void __launch_bounds__(256) join_probe_distinct_kernel(const __grid_constant__ KeyCols probe, const __grid_constant__ KeyCols build, int64_t n,
                                           const uint64_t* __restrict__ slots, uint32_t mask, bool nulls_equal,
                                           bool fast, int kind, unsigned long long* __restrict__ total, int32_t* __restrict__ left_map,
                                           int32_t* __restrict__ right_map, const unsigned long long* __restrict__ bloom, uint32_t bloom_mask,
                                           const int32_t* __restrict__ sel) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t base = warp * (32 * PI); base < n; base += nwarps * (32 * PI)) {
    int64_t src[PI];
    uint64_t kb[PI];
    uint32_t h[PI];
    bool live[PI];
    int32_t br[PI];
    // selection vector: probe row rr of the (virtual) filtered batch is row sel[rr] of the batch itself; the left map then
    // carries ORIGINAL row ids, so the payload gather reads the unfiltered batch and no filtered copy ever exists
#pragma unroll
    for (int j = 0; j < PI; j++) {
      const int64_t rr = base + j * 32 + lane;
      src[j] = rr < n ? (sel ? (int64_t)sel[rr] : rr) : -1;
    }
#pragma unroll
    for (int j = 0; j < PI; j++) {
      br[j] = INT32_MIN; kb[j] = 0; h[j] = 0;
      live[j] = src[j] >= 0 && ((nulls_equal && !fast) || !any_null_key(probe, src[j]));  // fast: the build side holds no NULL keys, so a NULL probe key matches nothing
      if (live[j]) { if (fast) { kb[j] = pack_join_key(probe, src[j]); h[j] = hash_packed(kb[j]); } else h[j] = row_hash(probe, src[j]); }
    }
    if (bloom) {
      unsigned long long w[PI];
#pragma unroll
      for (int j = 0; j < PI; j++) {
        uint32_t wi = 0; unsigned long long bits = 0;
        bloom_of(h[j], bloom_mask, wi, bits);
        w[j] = live[j] ? __ldg(&bloom[wi]) : 0ull;
        kb[j] = fast ? kb[j] : 0;
        live[j] = live[j] && (w[j] & bits) == bits;
      }
    }
#pragma unroll
    for (int j = 0; j < PI; j++) {
      if (!live[j]) continue;
      uint32_t idx = h[j] & mask;
      while (true) {
        uint64_t ek;
        const uint64_t e = join_entry(slots, idx, fast, ek);
        if (e == JSLOT_EMPTY) break;
        if ((uint32_t)(e >> 32) == h[j]) {
          const bool eq = fast ? (ek == kb[j]) : rows_equal(probe, src[j], build, (int32_t)(uint32_t)e, nulls_equal);
          if (eq) { br[j] = (int32_t)(uint32_t)e; break; }
        }
        idx = (idx + 1) & mask;
      }
    }
    if (kind == B2_JOIN_LEFT_OUTER) {
#pragma unroll
      for (int j = 0; j < PI; j++) {
        const int64_t rr = base + j * 32 + lane;
        if (rr < n) { left_map[rr] = (int32_t)src[j]; right_map[rr] = br[j]; }
      }
    } else {
      uint32_t ball[PI];
      uint32_t hits = 0;
#pragma unroll
      for (int j = 0; j < PI; j++) { ball[j] = __ballot_sync(0xffffffffu, br[j] != INT32_MIN); hits += __popc(ball[j]); }
      unsigned long long o = 0;
      if (lane == 0 && hits) o = atomicAdd(total, (unsigned long long)hits);
      o = __shfl_sync(0xffffffffu, o, 0);
#pragma unroll
      for (int j = 0; j < PI; j++) {
        if (br[j] != INT32_MIN) { const unsigned long long at = o + __popc(ball[j] & ((1u << lane) - 1u)); left_map[at] = (int32_t)src[j]; right_map[at] = br[j]; }
        o += __popc(ball[j]);
      }
    }
  }
}

// Keeps a region (the Bloom filter) in the persisting part of L2 while probe kernels stream the probe columns past it
// (cudaStreamAttributeAccessPolicyWindow); off unless B2_JOIN_BLOOM_PERSIST is set.
struct L2Persist {
  bool on = false;
  L2Persist(void* p, size_t bytes) {
    static const bool enabled = getenv("B2_JOIN_BLOOM_PERSIST") != nullptr;
    if (!enabled || !p || !bytes) return;
    static size_t max_window = 0, carve = 0;
    static bool init = false;
    if (!init) {
      init = true;
      int dev = 0; cudaGetDevice(&dev);
      cudaDeviceProp prop; cudaGetDeviceProperties(&prop, dev);
      carve = std::min<size_t>((size_t)prop.persistingL2CacheMaxSize, (size_t)48 << 20);
      max_window = (size_t)prop.accessPolicyMaxWindowSize;
      if (carve) cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, carve);
    }
    if (!carve || !max_window) return;
    cudaStreamAttrValue a; memset(&a, 0, sizeof(a));
    a.accessPolicyWindow.base_ptr = p;
    a.accessPolicyWindow.num_bytes = std::min(bytes, max_window);
    a.accessPolicyWindow.hitRatio = (float)std::min(1.0, (double)carve / (double)a.accessPolicyWindow.num_bytes);
    a.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    a.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    on = cudaStreamSetAttribute(stream(), cudaStreamAttributeAccessPolicyWindow, &a) == cudaSuccess;
    if (!on) cudaGetLastError();
  }
  ~L2Persist() {
    if (!on) return;
    cudaStreamAttrValue a; memset(&a, 0, sizeof(a));
    a.accessPolicyWindow.num_bytes = 0;
    cudaStreamSetAttribute(stream(), cudaStreamAttributeAccessPolicyWindow, &a);
  }
};

// The commonest probe of all — INNER join against a distinct build side on ONE integer key column without NULLs (every
// FK -> PK join of TPC-H) — without the generic row machinery (KeyCols loops, validity, runtime join kind): ~6x fewer
// instructions per row than join_probe_distinct_kernel.  Each warp takes 256 rows at a time:
//   1. row ids (through the selection vector), keys, Bloom words: 8 independent loads per lane at each step;
//   2. the rows that pass the filter (10-20 % in q3) are compacted into a per-warp queue in shared memory, so the random
//      HBM accesses into the table are issued by FULL warps in one or two rounds (the generic kernel walked its 8 row
//      slots one after the other, each round with 2-3 live lanes paying a full memory latency);
//   3. one output reservation (atomic) per round.
//   MODE 0: every row of the batch; 1: the rows of a selection vector (a filter below the join emitted row ids);
//   2: the filter itself is evaluated here (a "simple" predicate, simplefilter.cuh): the predicate columns of the 256 rows are
//      loaded first, rows that fail are dropped before their key is fetched, and no selection vector is ever written or read.
// (Tried and measured slower, 4.96 -> 5.47 ms per q3 step: evict-first loads of the selection vector and the key column plus
// evict-last Bloom words.  Unlike part_scatter2 — hash.cu — nothing here is half-written and waiting in L2.)
constexpr int PQ = 8;
template <typename T>
__device__ __forceinline__ uint32_t pred_term_mask(const SimpleTerm& t, const int32_t (&r)[PQ]) {
  T v[PQ];
#pragma unroll
  for (int j = 0; j < PQ; j++) v[j] = r[j] >= 0 ? reinterpret_cast<const T*>(t.col)[r[j]] : T(0);
  const T lit = (T)t.lit;
  uint32_t pass = 0;
#pragma unroll
  for (int j = 0; j < PQ; j++) {
    const int c = v[j] < lit ? 1 : (v[j] == lit ? 2 : 4);
    pass |= (uint32_t)((t.truth & c) != 0) << j;
  }
  return pass;
}
template <typename K, int MODE>
__global__ void __launch_bounds__(256) join_probe_distinct1_kernel(const K* __restrict__ keys, const int32_t* __restrict__ sel, int64_t n,
                                                                   const uint64_t* __restrict__ slots, uint32_t mask,
                                                                   const unsigned long long* __restrict__ bloom, uint32_t bloom_mask,
                                                                   unsigned long long* __restrict__ total, int32_t* __restrict__ left_map,
                                                                   int32_t* __restrict__ right_map, const __grid_constant__ SimplePred sp,
                                                                   unsigned long long* __restrict__ npass) {
  constexpr bool SEL = MODE == 1;
  typedef typename std::make_unsigned<K>::type UK;
  __shared__ uint8_t s_q[8][32 * PQ];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const uint32_t lt = (1u << lane) - 1u;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t base = warp * (32 * PQ); base < n; base += nwarps * (32 * PQ)) {
    int32_t r[PQ];
#pragma unroll
    for (int j = 0; j < PQ; j++) {
      const int64_t rr = base + j * 32 + lane;
      r[j] = rr < n ? (SEL ? sel[rr] : (int32_t)rr) : -1;
    }
    if (MODE == 2) {   // the filter: PQ independent loads per lane and term
      uint32_t ok = 0xffu;
      for (int k = 0; k < sp.n; k++) {
        const SimpleTerm& t = sp.t[k];
        switch (t.width) {
          case 1: ok &= pred_term_mask<int8_t>(t, r); break;
          case 2: ok &= pred_term_mask<int16_t>(t, r); break;
          case 4: ok &= pred_term_mask<int32_t>(t, r); break;
          default: ok &= pred_term_mask<int64_t>(t, r); break;
        }
      }
      int cnt = 0;
#pragma unroll
      for (int j = 0; j < PQ; j++) { if (!((ok >> j) & 1u)) r[j] = -1; cnt += r[j] >= 0; }
      for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
      if (lane == 0 && cnt) atomicAdd(npass, (unsigned long long)cnt);
    }
    uint32_t pass = 0;
    if (bloom) {
      uint32_t h[PQ];
      unsigned long long wv[PQ];
#pragma unroll
      for (int j = 0; j < PQ; j++) h[j] = r[j] >= 0 ? hash_packed((uint64_t)(UK)keys[r[j]]) : 0u;
#pragma unroll
      for (int j = 0; j < PQ; j++) {
        uint32_t wi; unsigned long long bits;
        bloom_of(h[j], bloom_mask, wi, bits);
        wv[j] = r[j] >= 0 ? __ldg(&bloom[wi]) : 0ull;
      }
#pragma unroll
      for (int j = 0; j < PQ; j++) {
        uint32_t wi; unsigned long long bits;
        bloom_of(h[j], bloom_mask, wi, bits);
        pass |= (uint32_t)(r[j] >= 0 && (wv[j] & bits) == bits) << j;
      }
    } else {
#pragma unroll
      for (int j = 0; j < PQ; j++) pass |= (uint32_t)(r[j] >= 0) << j;
    }
    int qn = 0;
#pragma unroll
    for (int j = 0; j < PQ; j++) {
      const bool p = (pass >> j) & 1u;
      const uint32_t b = __ballot_sync(0xffffffffu, p);
      if (p) s_q[w][qn + __popc(b & lt)] = (uint8_t)(j * 32 + lane);
      qn += __popc(b);
    }
    __syncwarp();
    for (int q0 = 0; q0 < qn; q0 += 32) {
      const int q = q0 + lane;
      int32_t br = INT32_MIN, src = 0;
      if (q < qn) {
        const int64_t rr = base + s_q[w][q];
        src = SEL ? sel[rr] : (int32_t)rr;
        const uint64_t kb = (uint64_t)(UK)keys[src];
        const uint32_t hh = hash_packed(kb);
        uint32_t idx = hh & mask;
        while (true) {
          const ulonglong2 e = *reinterpret_cast<const ulonglong2*>(&slots[(size_t)idx << 1]);
          if (e.x == JSLOT_EMPTY) break;
          if ((uint32_t)(e.x >> 32) == hh && e.y == kb) { br = (int32_t)(uint32_t)e.x; break; }
          idx = (idx + 1) & mask;
        }
      }
      const bool hit = br != INT32_MIN;
      const uint32_t b = __ballot_sync(0xffffffffu, hit);
      if (b) {
        unsigned long long o = 0;
        if (lane == 0) o = atomicAdd(total, (unsigned long long)__popc(b));
        o = __shfl_sync(0xffffffffu, o, 0) + __popc(b & lt);
        if (hit) { left_map[o] = src; right_map[o] = br; }
      }
    }
    __syncwarp();   // the queue is rewritten by the next chunk
  }
}

// MODE 0: count matches per probe row; MODE 1: write pairs at offsets
template <int MODE>
__global__ void join_probe_kernel(const __grid_constant__ KeyCols probe, const __grid_constant__ KeyCols build, int64_t n,
                                  const uint64_t* __restrict__ slots, uint32_t mask, bool nulls_equal,
                                  bool fast, int kind, int32_t* __restrict__ counts, const int64_t* __restrict__ offsets,
                                  int32_t* __restrict__ left_map, int32_t* __restrict__ right_map,
                                  const unsigned long long* __restrict__ bloom, uint32_t bloom_mask, const int32_t* __restrict__ sel) {
  for (int64_t rr = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; rr < n; rr += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = sel ? (int64_t)sel[rr] : rr;   // see join_probe_distinct_kernel
    int32_t matches = 0;
    int64_t o = MODE == 1 ? offsets[rr] : 0;
    const bool semi_like = kind == B2_JOIN_LEFT_SEMI || kind == B2_JOIN_LEFT_ANTI;
    if ((nulls_equal && !fast) || !any_null_key(probe, r)) {  // fast: see join_probe_distinct_kernel
      uint64_t kb = 0;
      uint32_t h;
      if (fast) { kb = pack_join_key(probe, r); h = hash_packed(kb); } else h = row_hash(probe, r);
      uint32_t idx = h & mask;
      const bool maybe = bloom_may_contain(bloom, bloom_mask, h);
      while (maybe) {
        uint64_t ek;
        const uint64_t e = join_entry(slots, idx, fast, ek);
        if (e == JSLOT_EMPTY) break;
        if ((uint32_t)(e >> 32) == h) {
          const int32_t br = (int32_t)(uint32_t)e;
          if (fast ? (ek == kb) : rows_equal(probe, r, build, br, nulls_equal)) {
            if (MODE == 1 && !semi_like) { left_map[o + matches] = (int32_t)r; right_map[o + matches] = br; }
            matches++;
            if (semi_like) break;
          }
        }
        idx = (idx + 1) & mask;
      }
    }
    int32_t emit;
    switch (kind) {
      case B2_JOIN_INNER: emit = matches; break;
      case B2_JOIN_LEFT_OUTER: emit = matches > 0 ? matches : 1; break;
      case B2_JOIN_LEFT_SEMI: emit = matches > 0 ? 1 : 0; break;
      default: emit = matches > 0 ? 0 : 1; break;  // LEFT_ANTI
    }
    if (MODE == 0) counts[rr] = emit;
    else {
      if (semi_like) { if (emit) left_map[o] = (int32_t)r; }
      else if (kind == B2_JOIN_LEFT_OUTER && matches == 0) { left_map[o] = (int32_t)r; right_map[o] = INT32_MIN; }
    }
  }
}

// ---- full outer = left outer + the build rows no stream row matched (GpuHashJoin.scala full-join gather maps) -----
__global__ void mark_matched_kernel(const int32_t* __restrict__ right_map, int64_t n, uint8_t* __restrict__ matched) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int32_t r = right_map[i];
    if (r >= 0) matched[r] = 1;
  }
}
__global__ void unmatched_flags_kernel(const uint8_t* __restrict__ matched, int64_t nb, int32_t* __restrict__ flags) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nb; i += (int64_t)gridDim.x * blockDim.x) flags[i] = matched[i] ? 0 : 1;
}
__global__ void append_unmatched_kernel(const uint8_t* __restrict__ matched, const int32_t* __restrict__ pos, int64_t nb, int64_t base,
                                        int32_t* __restrict__ left_map, int32_t* __restrict__ right_map) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nb; i += (int64_t)gridDim.x * blockDim.x)
    if (!matched[i]) { left_map[base + pos[i]] = INT32_MIN; right_map[base + pos[i]] = (int32_t)i; }
}

// ---- mixed (conditional) joins: which stream rows own at least one pair that passes the join condition ------------------------
__global__ void mark_passing_kernel(const int32_t* __restrict__ left_map, const int8_t* __restrict__ pass, const uint32_t* __restrict__ pass_valid, int64_t n,
                                    int8_t* __restrict__ flags) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    if (pass[i] && row_valid(pass_valid, i)) flags[left_map[i]] = 1;
}
__global__ void invert_flags_kernel(int8_t* __restrict__ flags, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) flags[i] = flags[i] ? 0 : 1;
}
// BOOL8 column over the stream rows: 1 where some (stream row, build row) pair of the equi-join passed the condition
// (invert: 1 where none did)
Column* rows_with_passing_pair(const Column* left_map, const Column* pass, int64_t stream_rows, bool invert) {
  ColGuard flags(new_column(B2_BOOL8, 0, stream_rows, false));
  if (stream_rows) CUDA_CHECK(cudaMemsetAsync(flags.c->data.p, 0, (size_t)stream_rows, stream()));
  if (left_map->size) {
    mark_passing_kernel<<<grid_for(left_map->size, 256), 256, 0, stream()>>>(left_map->data.as<int32_t>(), pass->data.as<int8_t>(), pass->validity(), left_map->size,
                                                                              flags.c->data.as<int8_t>());
    count_launch();
  }
  if (invert && stream_rows) { invert_flags_kernel<<<grid_for(stream_rows, 256), 256, 0, stream()>>>(flags.c->data.as<int8_t>(), stream_rows); count_launch(); }
  CUDA_CHECK(cudaGetLastError());
  return flags.release();
}

static JoinTable* jt_from(b2_handle h) {
  if (!h) throw Error(B2_ERR_INVALID, "null hash table handle");
  return reinterpret_cast<JoinTable*>((intptr_t)h);
}

// GpuFilter directly below the stream side of an INNER FK -> PK join, fused INTO the probe when the predicate is of the simple
// shape (simplefilter.cuh) and the probe qualifies for join_probe_distinct1_kernel: the gather maps carry ORIGINAL row ids of
// `batch`, `npass_out` = rows that passed the filter (the filter node's numOutputRows).  false = not applicable, the caller
// takes the selection-vector path.
bool join_probe_pred(b2_handle ht, const Table* batch, int key_col, const Program* prog, Column** out_lm, Column** out_rm, int64_t* npass_out) {
  // OFF by default: measured on the q3 step the fused kernel is slower (probe 5.2 -> 11.2 ms, step 24.1 -> 27.5 ms).  The probe
  // is bound by latency per row SLOT, not by bytes, and without the selection vector it walks all 600 M rows with 46 % of
  // its lanes idle; the filter kernel's compaction is worth more than the 1.3 GB round trip of the row ids.  Kept (and
  // tested) behind B2_JOIN_PRED_FUSION for a version that compacts the passing rows inside the warp first.
  if (!getenv("B2_JOIN_PRED_FUSION") || getenv("B2_JOIN_NO_FAST_PROBE")) return false;
  JoinTable* jt = jt_from(ht);
  const int64_t n = batch->rows;
  if (!jt->distinct || !jt->fast || jt->key_idx.size() != 1 || n < (1 << 16) || n >= 0x7fffffffLL) return false;
  const Column* pc = batch->cols[key_col];
  const int pw = dtype_width(pc->dtype);
  if (pc->nullable() || is_float(pc->dtype) || pc->dtype == B2_STRING || !(pw == 4 || pw == 8)) return false;
  if (pc->dtype != jt->keys->cols[jt->key_idx[0]]->dtype) return false;
  SimplePred sp;
  if (!simple_pred_of(prog, batch, sp)) return false;
  ColGuard lm(new_column(B2_INT32, 0, n, false)), rm(new_column(B2_INT32, 0, n, false));
  DevBuf tot(16);
  CUDA_CHECK(cudaMemsetAsync(tot.p, 0, 16, stream()));
  {
    KernelTimer kt("join_probe_distinct1_kernel");
    const int grid = grid_for(n, 256);
    const uint64_t* sl = jt->slots.as<uint64_t>(); const uint32_t msk = (uint32_t)(jt->cap - 1);
    const unsigned long long* bl = jt->bloom.as<unsigned long long>();
    unsigned long long* tp = tot.as<unsigned long long>();
    if (pw == 8) join_probe_distinct1_kernel<int64_t, 2><<<grid, 256, 0, stream()>>>(pc->data.as<int64_t>(), nullptr, n, sl, msk, bl, jt->bloom_mask, tp,
                                                                                       lm.c->data.as<int32_t>(), rm.c->data.as<int32_t>(), sp, tp + 1);
    else join_probe_distinct1_kernel<int32_t, 2><<<grid, 256, 0, stream()>>>(pc->data.as<int32_t>(), nullptr, n, sl, msk, bl, jt->bloom_mask, tp,
                                                                              lm.c->data.as<int32_t>(), rm.c->data.as<int32_t>(), sp, tp + 1);
    CUDA_CHECK(cudaGetLastError());
    count_launch();
  }
  unsigned long long h[2] = {0, 0};
  d2h(h, tot.p, 2);
  sync();
  lm.c->size = (int64_t)h[0]; rm.c->size = (int64_t)h[0];
  *npass_out = (int64_t)h[1];
  *out_lm = lm.release(); *out_rm = rm.release();
  return true;
}

}  // namespace b2

using namespace b2;
extern "C" {

int b2_join_build(b2_handle build_keys_table, int32_t nulls_equal, b2_handle* out_hash_table) {
  B2_TRY
  Table* t = table_from(build_keys_table);
  B2_CHECK(!t->cols.empty() && (int)t->cols.size() <= MAX_KEYS, "join needs 1..8 key columns");
  std::unique_ptr<JoinTable> jt(new JoinTable());
  t->refs.fetch_add(1);
  jt->keys = t;
  jt->build_rows = t->rows;
  jt->nulls_equal = nulls_equal != 0;
  for (int i = 0; i < (int)t->cols.size(); i++) jt->key_idx.push_back(i);
  int64_t cap = 1024;
  while (cap < t->rows * 2) cap <<= 1;
  jt->cap = cap;
  {
    int kw = 0; bool fixed = true, nullable = false;
    for (auto* c : t->cols) { if (c->dtype == B2_STRING || c->dtype == B2_DECIMAL128) fixed = false; kw += dtype_width(c->dtype); nullable = nullable || c->nullable(); }
    jt->fast = fixed && kw <= 8 && !(jt->nulls_equal && nullable);
  }
  jt->slots = DevBuf((size_t)cap * (jt->fast ? 16 : 8));
  CUDA_CHECK(cudaMemsetAsync(jt->slots.p, 0xff, jt->slots.bytes, stream()));
  if (t->rows >= (1 << 18) && !getenv("B2_JOIN_NO_BLOOM")) {   // smaller tables are L2 resident themselves
    // bits per key: fewer bits = more false positives (each costs one random HBM slot read) but a smaller filter.  Sweep on
    // the q3 step (probe kernel, ms per step): 16 bits 5.08, 8 bits 6.07, 4 bits 7.71; pinning the filter in the persisting
    // part of L2 (B2_JOIN_BLOOM_PERSIST) 5.31 — the false-positive rate matters more than L2 residency
    static const int bits_per_key = getenv("B2_JOIN_BLOOM_BITS") ? std::max(2, atoi(getenv("B2_JOIN_BLOOM_BITS"))) : 16;
    int64_t words = 1 << 15;
    while (words * 64 < t->rows * bits_per_key && words < (8 << 20)) words <<= 1;   // at most 64 MB
    if (words * 64 >= t->rows * 6) {                                // below ~6 bits per key the filter stops paying
      jt->bloom = DevBuf((size_t)words * 8);
      // bits set per key: sweep on the q3 step (probe kernel ms): 16 bits/key k=2 5.12, k=3 4.91; 8 bits/key k=3 5.58, k=4 5.43; 32 bits/key k=2 6.06
      static const int bloom_k = getenv("B2_JOIN_BLOOM_K") ? std::min(4, std::max(2, atoi(getenv("B2_JOIN_BLOOM_K")))) : 3;
      jt->bloom_mask = (uint32_t)(words - 1) | ((uint32_t)(bloom_k - 2) << 28);
      CUDA_CHECK(cudaMemsetAsync(jt->bloom.p, 0, jt->bloom.bytes, stream()));
    }
  }
  DevBuf dups(4);
  CUDA_CHECK(cudaMemsetAsync(dups.p, 0, 4, stream()));
  if (t->rows) {
    KeyCols keys = key_cols_of(t, jt->key_idx.data(), (int)jt->key_idx.size());
    KernelTimer kt_join_build_kernel("join_build_kernel");
    join_build_kernel<<<grid_for(t->rows, 256), 256, 0, stream()>>>(keys, t->rows, jt->slots.as<uint64_t>(), (uint32_t)(cap - 1),
                                                                    jt->nulls_equal, jt->fast, dups.as<int32_t>(), jt->bloom.as<unsigned long long>(), jt->bloom_mask);
    CUDA_CHECK(cudaGetLastError());
    count_launch();
  }
  int32_t h_dups = 0;
  d2h(&h_dups, dups.p, 1);
  sync();
  jt->distinct = h_dups == 0;
  *out_hash_table = to_handle(jt.release());
  B2_CATCH
}

int b2_join_hash_table_close(b2_handle ht) {
  B2_TRY
  delete jt_from(ht);
  B2_CATCH
}

int b2_join_probe(b2_handle ht, b2_handle probe_keys_table, int32_t kind, b2_handle* out_left_map, b2_handle* out_right_map) {
  return b2_join_probe_sel(ht, probe_keys_table, 0, kind, out_left_map, out_right_map);
}

// probe through a selection vector: `selection` (INT32 row ids into probe_keys_table, ascending, e.g. from b2_filter_row_ids)
// names the stream rows that take part; the left gather map carries ORIGINAL row ids of probe_keys_table's batch
int b2_join_probe_sel(b2_handle ht, b2_handle probe_keys_table, b2_handle selection, int32_t kind, b2_handle* out_left_map, b2_handle* out_right_map) {
  B2_TRY
  JoinTable* jt = jt_from(ht);
  Table* pt = table_from(probe_keys_table);
  const int32_t* sel = nullptr;
  int64_t nsel = 0;
  if (selection) {
    Column* sc = col_from(selection);
    B2_CHECK(sc->dtype == B2_INT32, "selection vector must be INT32");
    B2_CHECK(kind != B2_JOIN_FULL_OUTER, "full outer join through a selection vector");
    sel = sc->data.as<int32_t>(); nsel = sc->size;
  }
  B2_CHECK(pt->cols.size() == jt->keys->cols.size(), "probe and build key counts differ");
  for (size_t i = 0; i < pt->cols.size(); i++)
    B2_CHECK(pt->cols[i]->dtype == jt->keys->cols[i]->dtype, "probe and build key dtypes differ");
  if (kind == B2_JOIN_FULL_OUTER) {
    // left outer maps, then one extra row (left = out of bounds -> NULLs) per build row that nothing matched
    B2_CHECK(out_right_map != nullptr, "a full outer join needs both gather maps");
    b2_handle hl = 0, hr = 0;
    int rc = b2_join_probe_sel(ht, probe_keys_table, 0, B2_JOIN_LEFT_OUTER, &hl, &hr);
    if (rc != B2_OK) return rc;
    ColGuard lo(col_from(hl)), ro(col_from(hr));
    const int64_t m = lo.c->size, nb = jt->keys->rows;
    DevBuf matched((size_t)std::max<int64_t>(nb, 1)), pos((size_t)(nb + 1) * 4);
    CUDA_CHECK(cudaMemsetAsync(matched.p, 0, matched.bytes, stream()));
    int32_t extra = 0;
    if (nb) {
      if (m) mark_matched_kernel<<<grid_for(m, 256), 256, 0, stream()>>>(ro.c->data.as<int32_t>(), m, matched.as<uint8_t>());
      unmatched_flags_kernel<<<grid_for(nb, 256), 256, 0, stream()>>>(matched.as<uint8_t>(), nb, pos.as<int32_t>());
      count_launch(2);
      exclusive_scan<int32_t, int32_t>(pos.as<int32_t>(), pos.as<int32_t>(), nb, true);
      d2h(&extra, pos.as<int32_t>() + nb, 1);
      sync();
    }
    if (m + extra > 0x7fffffffLL) throw Error(B2_ERR_SIZE_OVERFLOW, "join output exceeds 2^31-1 rows; split the stream batch");
    ColGuard lm(new_column(B2_INT32, 0, m + extra, false)), rm(new_column(B2_INT32, 0, m + extra, false));
    if (m) {
      CUDA_CHECK(cudaMemcpyAsync(lm.c->data.p, lo.c->data.p, (size_t)m * 4, cudaMemcpyDeviceToDevice, stream()));
      CUDA_CHECK(cudaMemcpyAsync(rm.c->data.p, ro.c->data.p, (size_t)m * 4, cudaMemcpyDeviceToDevice, stream()));
    }
    if (extra) {
      append_unmatched_kernel<<<grid_for(nb, 256), 256, 0, stream()>>>(matched.as<uint8_t>(), pos.as<int32_t>(), nb, m, lm.c->data.as<int32_t>(),
                                                                        rm.c->data.as<int32_t>());
      CUDA_CHECK(cudaGetLastError());
      count_launch();
    }
    *out_left_map = to_handle(lm.release());
    *out_right_map = to_handle(rm.release());
    return B2_OK;
  }
  B2_CHECK(kind >= B2_JOIN_INNER && kind <= B2_JOIN_LEFT_ANTI, "bad join kind");
  const int64_t n = sel ? nsel : pt->rows;
  const bool semi_like = kind == B2_JOIN_LEFT_SEMI || kind == B2_JOIN_LEFT_ANTI;
  KeyCols pk = key_cols_of(pt, jt->key_idx.data(), (int)jt->key_idx.size());
  KeyCols bk = key_cols_of(jt->keys, jt->key_idx.data(), (int)jt->key_idx.size());
  if (jt->distinct && (kind == B2_JOIN_INNER || kind == B2_JOIN_LEFT_OUTER)) {
    // innerDistinctJoinGatherMaps / leftDistinctJoinGatherMap fast path (GpuHashJoin.scala:1401-1428)
    ColGuard lm(new_column(B2_INT32, 0, n, false)), rm(new_column(B2_INT32, 0, n, false));
    int64_t matched = n;
    if (n) {
      DevBuf tot(8);
      CUDA_CHECK(cudaMemsetAsync(tot.p, 0, 8, stream()));
      L2Persist keep(jt->bloom.p, jt->bloom.bytes);
      const Column* pc = jt->key_idx.size() == 1 ? pt->cols[jt->key_idx[0]] : nullptr;
      const int pw = pc ? dtype_width(pc->dtype) : 0;
      if (kind == B2_JOIN_INNER && jt->fast && pc && !pc->nullable() && !is_float(pc->dtype) && pc->dtype != B2_STRING && (pw == 4 || pw == 8) &&
          n < 0x7fffffffLL && !getenv("B2_JOIN_NO_FAST_PROBE")) {
        KernelTimer kt("join_probe_distinct1_kernel");
        const int grid = grid_for(n, 256);
        const uint64_t* sl = jt->slots.as<uint64_t>(); const uint32_t msk = (uint32_t)(jt->cap - 1);
        const unsigned long long* bl = jt->bloom.as<unsigned long long>(); unsigned long long* tp = tot.as<unsigned long long>();
        int32_t* lp = lm.c->data.as<int32_t>(); int32_t* rp = rm.c->data.as<int32_t>();
        SimplePred none; memset(&none, 0, sizeof(none));
        if (pw == 8) {
          if (sel) join_probe_distinct1_kernel<int64_t, 1><<<grid, 256, 0, stream()>>>(pc->data.as<int64_t>(), sel, n, sl, msk, bl, jt->bloom_mask, tp, lp, rp, none, nullptr);
          else join_probe_distinct1_kernel<int64_t, 0><<<grid, 256, 0, stream()>>>(pc->data.as<int64_t>(), sel, n, sl, msk, bl, jt->bloom_mask, tp, lp, rp, none, nullptr);
        } else {
          if (sel) join_probe_distinct1_kernel<int32_t, 1><<<grid, 256, 0, stream()>>>(pc->data.as<int32_t>(), sel, n, sl, msk, bl, jt->bloom_mask, tp, lp, rp, none, nullptr);
          else join_probe_distinct1_kernel<int32_t, 0><<<grid, 256, 0, stream()>>>(pc->data.as<int32_t>(), sel, n, sl, msk, bl, jt->bloom_mask, tp, lp, rp, none, nullptr);
        }
        CUDA_CHECK(cudaGetLastError());
        count_launch();
      } else {
      KernelTimer kt("join_probe_distinct_kernel");
      join_probe_distinct_kernel<<<grid_for(n, 256), 256, 0, stream()>>>(pk, bk, n, jt->slots.as<uint64_t>(), (uint32_t)(jt->cap - 1),
                                                                          jt->nulls_equal, jt->fast, kind, tot.as<unsigned long long>(),
                                                                          lm.c->data.as<int32_t>(), rm.c->data.as<int32_t>(), jt->bloom.as<unsigned long long>(), jt->bloom_mask, sel);
      CUDA_CHECK(cudaGetLastError());
      count_launch();
      }
      if (kind == B2_JOIN_INNER) { unsigned long long h = 0; d2h(&h, tot.p, 1); sync(); matched = (int64_t)h; }
    }
    lm.c->size = matched; rm.c->size = matched;  // buffers stay sized for n rows
    *out_left_map = to_handle(lm.release());
    if (out_right_map) *out_right_map = to_handle(rm.release()); else col_release(rm.release());
    return B2_OK;
  }
  DevBuf counts((size_t)std::max<int64_t>(n, 1) * 4), offsets((size_t)(n + 1) * 8);
  int64_t total = 0;
  if (n) {
    KernelTimer kt_join_probe_count_kernel("join_probe_count_kernel");
    join_probe_kernel<0><<<grid_for(n, 256), 256, 0, stream()>>>(pk, bk, n, jt->slots.as<uint64_t>(), (uint32_t)(jt->cap - 1), jt->nulls_equal, jt->fast, kind,
                                                                  counts.as<int32_t>(), nullptr, nullptr, nullptr, jt->bloom.as<unsigned long long>(), jt->bloom_mask, sel);
    CUDA_CHECK(cudaGetLastError());
    count_launch();
    exclusive_scan<int32_t, int64_t>(counts.as<int32_t>(), offsets.as<int64_t>(), n, true);
    d2h(&total, offsets.as<int64_t>() + n, 1);
    sync();
  }
  // the reference splits the stream batch when the maps would pass the batch target
  // (AbstractGpuJoinIterator.scala:235-250); the hard limit here is the int32 row index
  if (total > 0x7fffffffLL) throw Error(B2_ERR_SIZE_OVERFLOW, "join output exceeds 2^31-1 rows; split the stream batch");
  ColGuard lm(new_column(B2_INT32, 0, total, false));
  ColGuard rm(semi_like ? nullptr : new_column(B2_INT32, 0, total, false));
  if (n && total) {
    KernelTimer kt_join_probe_write_kernel("join_probe_write_kernel");
    join_probe_kernel<1><<<grid_for(n, 256), 256, 0, stream()>>>(pk, bk, n, jt->slots.as<uint64_t>(), (uint32_t)(jt->cap - 1), jt->nulls_equal, jt->fast, kind,
                                                                  nullptr, offsets.as<int64_t>(), lm.c->data.as<int32_t>(),
                                                                  semi_like ? nullptr : rm.c->data.as<int32_t>(), jt->bloom.as<unsigned long long>(), jt->bloom_mask, sel);
    CUDA_CHECK(cudaGetLastError());
    count_launch();
  }
  *out_left_map = to_handle(lm.release());
  if (out_right_map) *out_right_map = semi_like ? 0 : to_handle(rm.release());
  B2_CATCH
}

}  // extern "C"
