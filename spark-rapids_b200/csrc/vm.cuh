// vm.cuh — the fused expression evaluator: a register machine whose registers are typed column
// slices in shared memory.  One CTA owns a tile of VM_TILE rows; thread t owns rows t + j*VM_NT,
// so every register access is thread-private (no barriers between instructions) and bank-conflict
// free, while every global column access is fully coalesced (a warp reads 32 consecutive rows).
//
// Replaces the reference's one-cudf-kernel-per-expression-node evaluation with full intermediate
// columns (GpuExpressions.scala:397-413 CudfBinaryExpression.doColumnar): here the
// whole bound expression list of a GpuProjectExec / GpuFilterExec / aggregate pre-step is ONE
// kernel, intermediates never leave the SM.
#pragma once
#include "common.cuh"

namespace b2 {

constexpr int VM_NT = 256;           // threads per CTA
constexpr int VM_K = 4;              // rows per thread per tile
constexpr int VM_TILE = VM_NT * VM_K;
constexpr int VM_MAX_REGS = 64;
constexpr int VM_MAX_COLS = 64;
constexpr int VM_MAX_OUTS = 32;

enum MT : uint8_t { MT_I8 = 0, MT_I16 = 1, MT_I32 = 2, MT_I64 = 3, MT_I128 = 4, MT_F32 = 5, MT_F64 = 6 };
__host__ __device__ inline int mt_width(int mt) {
  switch (mt) { case MT_I8: return 1; case MT_I16: return 2; case MT_I32: case MT_F32: return 4;
                case MT_I64: case MT_F64: return 8; default: return 16; }
}

enum OK : uint8_t { OK_NONE = 0, OK_REG = 1, OK_COL = 2, OK_LIT = 3 };

enum VOP : uint8_t {
  V_ADD = 1, V_SUB, V_MUL, V_DIV, V_MOD, V_PMOD, V_NEG, V_ABS,
  V_EQ, V_NE, V_LT, V_LE, V_GT, V_GE, V_EQNS,
  V_AND, V_OR, V_NOT,
  V_ISNULL, V_ISNOTNULL, V_COALESCE, V_IF,
  V_CAST,        // mt = source, mt2 = destination machine type
  V_RESCALE_UP,  // integer * 10^aux with overflow -> null (decimal scale increase); mt=mt2 width
  V_RESCALE_DOWN,// integer / 10^aux HALF_UP (decimal scale decrease)
  V_CHECK_PREC,  // |x| >= 10^aux -> null  (CheckOverflow / GpuCheckOverflow)
  V_MULDEC,      // 128x128 -> 256-bit product, / 10^aux HALF_UP, overflow -> null
  V_DEC2F64,     // decimal (mt) -> double, / 10^aux
  V_NORM_NAN_ZERO,
  V_YEAR,
  V_MOV
};

struct alignas(16) VMOperand {
  int64_t lo, hi;    // OK_LIT value (first so that 16-byte literal loads are aligned)
  int32_t idx;       // register or input column index
  uint8_t kind;      // OK
  uint8_t nullable;  // may carry nulls
  uint8_t lit_null;  // OK_LIT: literal is NULL
  uint8_t pad;
};

struct alignas(16) VMInstr {
  uint8_t op, mt, mt2, dst_nullable;
  int32_t dst;
  int32_t aux;
  int32_t pad;
  VMOperand a, b, c;
};

struct VMReg { int32_t off; int32_t voff; };  // byte offsets inside the tile's shared memory

struct VMProgramHeader {
  int32_t ninstr, nregs, ncols, nouts;
  int32_t smem_bytes;
  VMReg regs[VM_MAX_REGS];
  VMOperand outs[VM_MAX_OUTS];
  uint8_t out_mt[VM_MAX_OUTS];
};

struct VMInputs {
  const void* data[VM_MAX_COLS];
  const uint32_t* valid[VM_MAX_COLS];
};

#ifdef __CUDACC__
// ------------------------------------------------------------------------------------------------

__device__ __forceinline__ i128 pow10_i128(int e) {
  // 10^e for 0 <= e <= 38
  const unsigned long long p19 = 10000000000000000000ull;  // 10^19
  unsigned long long small = 1;
  int r = e >= 19 ? e - 19 : e;
  for (int i = 0; i < r; i++) small *= 10ull;
  return e >= 19 ? (i128)((u128)small * (u128)p19) : (i128)small;
}

// resolved per instruction, uniform across the CTA
struct Opnd {
  const char* base;  // generic pointer: shared register slice, global column (pre-offset to tile) or literal
  int stride;        // bytes per row (0 for a literal)
  int vkind;         // 0 always valid, 1 byte-per-row (shared), 2 bitmask (global), 3 always null
  const uint8_t* vbytes;
  const uint32_t* vbits;
};

struct VMCtx {
  const VMProgramHeader* hdr;
  const VMInputs* in;
  char* smem;
  int64_t tile_base;  // first global row of this tile
  int64_t nrows;
};

__device__ __forceinline__ Opnd resolve(const VMCtx& cx, const VMOperand& o, int width) {
  Opnd r;
  r.vbytes = nullptr; r.vbits = nullptr;
  if (o.kind == OK_REG) {
    r.base = cx.smem + cx.hdr->regs[o.idx].off;
    r.stride = width;
    r.vkind = o.nullable ? 1 : 0;
    r.vbytes = reinterpret_cast<const uint8_t*>(cx.smem + cx.hdr->regs[o.idx].voff);
  } else if (o.kind == OK_COL) {
    r.base = reinterpret_cast<const char*>(cx.in->data[o.idx]) + cx.tile_base * width;
    r.stride = width;
    const uint32_t* vb = cx.in->valid[o.idx];
    r.vkind = (o.nullable && vb) ? 2 : 0;
    r.vbits = vb;
  } else {
    r.base = reinterpret_cast<const char*>(&o.lo);
    r.stride = 0;
    r.vkind = o.lit_null ? 3 : 0;
  }
  return r;
}

template <typename T>
__device__ __forceinline__ T opnd_ld(const Opnd& o, int i) {
  return *reinterpret_cast<const T*>(o.base + (size_t)i * o.stride);
}
__device__ __forceinline__ bool opnd_valid(const Opnd& o, int i, int64_t g) {
  switch (o.vkind) {
    case 0: return true;
    case 1: return o.vbytes[i] != 0;
    case 2: return bit_get(o.vbits, g);
    default: return false;
  }
}

struct Dst {
  char* base; int stride; uint8_t* vbytes; bool nullable;
};
__device__ __forceinline__ Dst resolve_dst(const VMCtx& cx, const VMInstr& ins, int width) {
  Dst d;
  d.base = cx.smem + cx.hdr->regs[ins.dst].off;
  d.stride = width;
  d.vbytes = reinterpret_cast<uint8_t*>(cx.smem + cx.hdr->regs[ins.dst].voff);
  d.nullable = ins.dst_nullable;
  return d;
}
template <typename T>
__device__ __forceinline__ void dst_st(const Dst& d, int i, T v, bool valid) {
  *reinterpret_cast<T*>(d.base + (size_t)i * d.stride) = v;
  if (d.nullable) d.vbytes[i] = valid ? 1 : 0;
}

#define VM_ROWS(i, g)                                         \
  _Pragma("unroll") for (int _j = 0; _j < VM_K; _j++)         \
    if (const int i = threadIdx.x + _j * VM_NT; true)         \
      if (const int64_t g = cx.tile_base + i; g < cx.nrows)

template <typename T> struct UnsignedOf { typedef T type; };
template <> struct UnsignedOf<int8_t> { typedef uint8_t type; };
template <> struct UnsignedOf<int16_t> { typedef uint16_t type; };
template <> struct UnsignedOf<int32_t> { typedef uint32_t type; };
template <> struct UnsignedOf<int64_t> { typedef uint64_t type; };
template <> struct UnsignedOf<i128> { typedef u128 type; };
template <typename T> struct IsFloat { static const bool v = false; };
template <> struct IsFloat<float> { static const bool v = true; };
template <> struct IsFloat<double> { static const bool v = true; };

// Spark comparison semantics (predicates.scala:155-331): NaN == NaN, NaN greater than all, -0.0 == 0.0
template <typename T>
__device__ __forceinline__ int cmp3(T a, T b) {
  if constexpr (IsFloat<T>::v) {
    bool an = a != a, bn = b != b;
    if (an || bn) return an == bn ? 0 : (an ? 1 : -1);
  }
  return a < b ? -1 : (a > b ? 1 : 0);
}

template <typename T>
__device__ __forceinline__ void vm_arith(const VMCtx& cx, const VMInstr& ins) {
  typedef typename UnsignedOf<T>::type U;
  Opnd a = resolve(cx, ins.a, sizeof(T)), b = resolve(cx, ins.b, sizeof(T));
  Dst d = resolve_dst(cx, ins, sizeof(T));
  const int op = ins.op;
  VM_ROWS(i, g) {
    T x = opnd_ld<T>(a, i), y = opnd_ld<T>(b, i);
    bool v = opnd_valid(a, i, g) && opnd_valid(b, i, g);
    T r = x;
    if constexpr (IsFloat<T>::v) {
      // arithmetic.scala:309-340 — IEEE; Divide/Remainder by zero -> NULL (Spark non-ANSI)
      switch (op) {
        case V_ADD: r = x + y; break;
        case V_SUB: r = x - y; break;
        case V_MUL: r = x * y; break;
        case V_DIV: if (y == (T)0) { v = false; r = 0; } else r = x / y; break;
        default: if (y == (T)0) { v = false; r = 0; } else {
          r = (T)fmod((double)x, (double)y);
          if (op == V_PMOD && r != (T)0 && ((r < 0) != (y < 0))) r += y; } break;
      }
    } else {
      switch (op) {  // integer: two's-complement wrap (arithmetic.scala:38-75, non-ANSI)
        case V_ADD: r = (T)((U)x + (U)y); break;
        case V_SUB: r = (T)((U)x - (U)y); break;
        case V_MUL: r = (T)((U)x * (U)y); break;
        default:
          if (y == (T)0) { v = false; r = 0; }
          else if (y == (T)-1) { r = (op == V_DIV) ? (T)((U)0 - (U)x) : (T)0; }
          else if (op == V_DIV) r = x / y;
          else { r = x % y; if (op == V_PMOD && r != 0 && ((r < 0) != (y < 0))) r += y; }
          break;
      }
    }
    dst_st<T>(d, i, r, v);
  }
}

// 128-bit decimal add/sub with overflow -> null (arithmetic.scala:78-125)
__device__ __forceinline__ void vm_arith128(const VMCtx& cx, const VMInstr& ins) {
  Opnd a = resolve(cx, ins.a, 16), b = resolve(cx, ins.b, 16);
  Dst d = resolve_dst(cx, ins, 16);
  const int op = ins.op;
  VM_ROWS(i, g) {
    i128 x = opnd_ld<i128>(a, i), y = opnd_ld<i128>(b, i);
    bool v = opnd_valid(a, i, g) && opnd_valid(b, i, g);
    i128 r;
    if (op == V_ADD) { r = (i128)((u128)x + (u128)y); if (((x ^ r) & (y ^ r)) < 0) v = false; }
    else if (op == V_SUB) { r = (i128)((u128)x - (u128)y); if (((x ^ y) & (x ^ r)) < 0) v = false; }
    else { r = (i128)((u128)x * (u128)y); }  // V_MUL: caller guarantees it fits (p1+p2+1 <= 38)
    dst_st<i128>(d, i, r, v);
  }
}

template <typename T>
__device__ __forceinline__ void vm_compare(const VMCtx& cx, const VMInstr& ins) {
  Opnd a = resolve(cx, ins.a, sizeof(T)), b = resolve(cx, ins.b, sizeof(T));
  Dst d = resolve_dst(cx, ins, 1);
  const int op = ins.op;
  VM_ROWS(i, g) {
    T x = opnd_ld<T>(a, i), y = opnd_ld<T>(b, i);
    bool va = opnd_valid(a, i, g), vb = opnd_valid(b, i, g);
    int c = cmp3<T>(x, y);
    bool r, v = va && vb;
    switch (op) {
      case V_EQ: r = c == 0; break;
      case V_NE: r = c != 0; break;
      case V_LT: r = c < 0; break;
      case V_LE: r = c <= 0; break;
      case V_GT: r = c > 0; break;
      case V_GE: r = c >= 0; break;
      default:   r = (va && vb) ? (c == 0) : (va == vb); v = true; break;  // <=> EqualNullSafe
    }
    dst_st<int8_t>(d, i, (int8_t)(r && (v || op == V_EQNS)), v);
  }
}

__device__ __forceinline__ void vm_logic(const VMCtx& cx, const VMInstr& ins) {
  Opnd a = resolve(cx, ins.a, 1);
  Dst d = resolve_dst(cx, ins, 1);
  if (ins.op == V_NOT) {
    VM_ROWS(i, g) { bool v = opnd_valid(a, i, g); dst_st<int8_t>(d, i, (int8_t)(v && !opnd_ld<int8_t>(a, i)), v); }
    return;
  }
  Opnd b = resolve(cx, ins.b, 1);
  const bool is_and = ins.op == V_AND;
  VM_ROWS(i, g) {  // Kleene logic, predicates.scala:54-153 (NULL_LOGICAL_AND / NULL_LOGICAL_OR)
    bool va = opnd_valid(a, i, g), vb = opnd_valid(b, i, g);
    bool x = va && opnd_ld<int8_t>(a, i), y = vb && opnd_ld<int8_t>(b, i);
    bool r, v;
    if (is_and) { bool fa = va && !x, fb = vb && !y; r = x && y; v = (va && vb) || fa || fb; }
    else { r = x || y; v = (va && vb) || x || y; }
    dst_st<int8_t>(d, i, (int8_t)(r && v), v);
  }
}

template <typename T>
__device__ __forceinline__ void vm_select(const VMCtx& cx, const VMInstr& ins) {
  Dst d = resolve_dst(cx, ins, sizeof(T));
  if (ins.op == V_COALESCE) {
    Opnd a = resolve(cx, ins.a, sizeof(T)), b = resolve(cx, ins.b, sizeof(T));
    VM_ROWS(i, g) {
      bool va = opnd_valid(a, i, g);
      T r = va ? opnd_ld<T>(a, i) : opnd_ld<T>(b, i);
      dst_st<T>(d, i, r, va || opnd_valid(b, i, g));
    }
  } else if (ins.op == V_IF) {  // conditionalExpressions.scala GpuIf: null predicate takes the else branch
    Opnd p = resolve(cx, ins.a, 1), a = resolve(cx, ins.b, sizeof(T)), b = resolve(cx, ins.c, sizeof(T));
    VM_ROWS(i, g) {
      bool t = opnd_valid(p, i, g) && opnd_ld<int8_t>(p, i);
      T r = t ? opnd_ld<T>(a, i) : opnd_ld<T>(b, i);
      dst_st<T>(d, i, r, t ? opnd_valid(a, i, g) : opnd_valid(b, i, g));
    }
  } else if (ins.op == V_MOV) {
    Opnd a = resolve(cx, ins.a, sizeof(T));
    VM_ROWS(i, g) { dst_st<T>(d, i, opnd_ld<T>(a, i), opnd_valid(a, i, g)); }
  } else if (ins.op == V_NEG || ins.op == V_ABS) {
    typedef typename UnsignedOf<T>::type U;
    Opnd a = resolve(cx, ins.a, sizeof(T));
    const bool neg = ins.op == V_NEG;
    VM_ROWS(i, g) {
      T x = opnd_ld<T>(a, i);
      T r;
      if constexpr (IsFloat<T>::v) r = neg ? -x : (x < 0 || (x == 0 && 1 / (double)x < 0) ? -x : x);
      else r = (neg || x < 0) ? (T)((U)0 - (U)x) : x;
      dst_st<T>(d, i, r, opnd_valid(a, i, g));
    }
  }
}

// GpuCast.scala:295 doCast, numeric subset.  Integral narrowing wraps (Java semantics); float ->
// integral follows Java (NaN -> 0, saturating) via int/long then narrows.
// correctly rounded signed 128-bit -> double (keep 64 significant bits + sticky, then scale)
__device__ __forceinline__ double i128_to_double(i128 x) {
  bool neg = x < 0;
  u128 m = neg ? (u128)0 - (u128)x : (u128)x;
  uint64_t hi = (uint64_t)(m >> 64), lo = (uint64_t)m;
  double r;
  if (hi == 0) r = (double)lo;
  else {
    int lz = __clzll((long long)hi);
    int sh = 64 - lz;  // bits shifted out of the low word
    uint64_t top = (uint64_t)(m >> sh);
    bool sticky = (m & ((((u128)1) << sh) - 1)) != 0;
    top |= sticky ? 1ull : 0ull;
    r = ldexp((double)top, sh);
  }
  return neg ? -r : r;
}
template <typename S, typename D> struct CastVia { __device__ static __forceinline__ D f(S x) { return (D)x; } };
template <> struct CastVia<i128, double> { __device__ static __forceinline__ double f(i128 x) { return i128_to_double(x); } };
template <> struct CastVia<i128, float> { __device__ static __forceinline__ float f(i128 x) { return (float)i128_to_double(x); } };

template <typename S, typename D>
__device__ __forceinline__ D cast_val(S x) {
  if constexpr (IsFloat<S>::v && !IsFloat<D>::v) {
    if constexpr (sizeof(D) >= 8) {
      if (x != x) return (D)0;
      if (x >= (S)9223372036854775807.0) return (D)0x7fffffffffffffffLL;
      if (x <= (S)-9223372036854775808.0) return (D)(-0x7fffffffffffffffLL - 1);
      return (D)(long long)x;
    } else {
      int v;
      if (x != x) v = 0;
      else if (x >= (S)2147483647.0) v = 0x7fffffff;
      else if (x <= (S)-2147483648.0) v = -0x7fffffff - 1;
      else v = (int)x;
      return (D)v;
    }
  } else {
    return CastVia<S, D>::f(x);
  }
}
template <typename S, typename D>
__device__ __forceinline__ void vm_cast2(const VMCtx& cx, const VMInstr& ins) {
  Opnd a = resolve(cx, ins.a, sizeof(S));
  Dst d = resolve_dst(cx, ins, sizeof(D));
  VM_ROWS(i, g) { dst_st<D>(d, i, cast_val<S, D>(opnd_ld<S>(a, i)), opnd_valid(a, i, g)); }
}
template <typename S>
__device__ __forceinline__ void vm_cast1(const VMCtx& cx, const VMInstr& ins) {
  switch (ins.mt2) {
    case MT_I8: vm_cast2<S, int8_t>(cx, ins); break;
    case MT_I16: vm_cast2<S, int16_t>(cx, ins); break;
    case MT_I32: vm_cast2<S, int32_t>(cx, ins); break;
    case MT_I64: vm_cast2<S, int64_t>(cx, ins); break;
    case MT_I128: vm_cast2<S, i128>(cx, ins); break;
    case MT_F32: vm_cast2<S, float>(cx, ins); break;
    default: vm_cast2<S, double>(cx, ins); break;
  }
}

// divide a 256-bit magnitude (4 x u64 little endian) by a u64, returning the remainder
__device__ __forceinline__ uint64_t div256_u64(uint64_t q[4], uint64_t dv) {
  u128 rem = 0;
  for (int k = 3; k >= 0; k--) {
    u128 cur = (rem << 64) | q[k];
    q[k] = (uint64_t)(cur / dv);
    rem = cur % dv;
  }
  return (uint64_t)rem;
}

template <typename T>
__device__ __forceinline__ void vm_decimal(const VMCtx& cx, const VMInstr& ins) {
  Opnd a = resolve(cx, ins.a, sizeof(T));
  Dst d = resolve_dst(cx, ins, sizeof(T));
  const i128 p = pow10_i128(ins.aux);
  if (ins.op == V_RESCALE_UP) {
    VM_ROWS(i, g) {
      i128 x = (i128)opnd_ld<T>(a, i);
      bool v = opnd_valid(a, i, g);
      // overflow if |x| * 10^aux does not fit T
      const i128 maxv = (i128)((((u128)1) << (8 * sizeof(T) - 1)) - 1);
      const i128 lim = maxv / p;
      i128 r = (i128)((u128)x * (u128)p);
      if (x > lim || x < -lim) v = false;
      dst_st<T>(d, i, (T)r, v);
    }
  } else if (ins.op == V_RESCALE_DOWN) {  // HALF_UP (away from zero), as BigDecimal.setScale
    VM_ROWS(i, g) {
      i128 x = (i128)opnd_ld<T>(a, i);
      bool neg = x < 0;
      u128 m = neg ? (u128)0 - (u128)x : (u128)x;
      u128 q = m / (u128)p, r = m % (u128)p;
      if (r * 2 >= (u128)p) q += 1;
      i128 res = neg ? -(i128)q : (i128)q;
      dst_st<T>(d, i, (T)res, opnd_valid(a, i, g));
    }
  } else {  // V_CHECK_PREC
    VM_ROWS(i, g) {
      T x = opnd_ld<T>(a, i);
      bool v = opnd_valid(a, i, g);
      i128 xx = (i128)x;
      if (xx >= p || xx <= -p) v = false;
      dst_st<T>(d, i, x, v);
    }
  }
}

// DecimalUtils.multiply128 (arithmetic.scala:470-512 longMultiply): exact 256-bit product,
// HALF_UP to the result scale, NULL when the result needs more than 38 digits.
__device__ __forceinline__ void vm_muldec(const VMCtx& cx, const VMInstr& ins) {
  Opnd a = resolve(cx, ins.a, 16), b = resolve(cx, ins.b, 16);
  Dst d = resolve_dst(cx, ins, 16);
  const int k = ins.aux;
  VM_ROWS(i, g) {
    i128 x = opnd_ld<i128>(a, i), y = opnd_ld<i128>(b, i);
    bool v = opnd_valid(a, i, g) && opnd_valid(b, i, g);
    bool neg = (x < 0) != (y < 0);
    u128 mx = x < 0 ? (u128)0 - (u128)x : (u128)x, my = y < 0 ? (u128)0 - (u128)y : (u128)y;
    uint64_t x0 = (uint64_t)mx, x1 = (uint64_t)(mx >> 64), y0 = (uint64_t)my, y1 = (uint64_t)(my >> 64);
    uint64_t q[4];
    u128 p00 = (u128)x0 * y0, p01 = (u128)x0 * y1, p10 = (u128)x1 * y0, p11 = (u128)x1 * y1;
    q[0] = (uint64_t)p00;
    u128 mid = (p00 >> 64) + (uint64_t)p01 + (uint64_t)p10;
    q[1] = (uint64_t)mid;
    u128 hi = (mid >> 64) + (p01 >> 64) + (p10 >> 64) + (uint64_t)p11;
    q[2] = (uint64_t)hi;
    q[3] = (uint64_t)((hi >> 64) + (p11 >> 64));
    if (k > 0) {
      // divide by 10^k in up to three u64 steps, tracking whether remainder*2 >= 10^k
      int k1 = k > 19 ? 19 : k;
      uint64_t d1 = 1; for (int t = 0; t < k1; t++) d1 *= 10ull;
      uint64_t r1 = div256_u64(q, d1);
      int k2 = k - k1;
      if (k2 > 0) {
        uint64_t d2 = 1; for (int t = 0; t < k2; t++) d2 *= 10ull;
        uint64_t r2 = div256_u64(q, d2);
        // total remainder = r2*d1 + r1 vs (d1*d2)/2
        u128 rem = (u128)r2 * d1 + r1, half = ((u128)d1 * d2) / 2;
        if (rem >= half) { for (int t = 0; t < 4; t++) { if (++q[t] != 0) break; } }
      } else {
        if ((u128)r1 * 2 >= (u128)d1) { for (int t = 0; t < 4; t++) { if (++q[t] != 0) break; } }
      }
    }
    u128 mag = ((u128)q[1] << 64) | q[0];
    if (q[2] || q[3] || mag >= (u128)pow10_i128(38)) v = false;
    i128 r = neg ? -(i128)mag : (i128)mag;
    dst_st<i128>(d, i, r, v);
  }
}

template <typename T>
__device__ __forceinline__ void vm_dec2f64(const VMCtx& cx, const VMInstr& ins) {
  Opnd a = resolve(cx, ins.a, sizeof(T));
  Dst d = resolve_dst(cx, ins, 8);
  double dv = 1.0; for (int t = 0; t < ins.aux; t++) dv *= 10.0;
  VM_ROWS(i, g) { dst_st<double>(d, i, CastVia<T, double>::f(opnd_ld<T>(a, i)) / dv, opnd_valid(a, i, g)); }
}

template <typename T>
__device__ __forceinline__ void vm_normnz(const VMCtx& cx, const VMInstr& ins) {
  Opnd a = resolve(cx, ins.a, sizeof(T));
  Dst d = resolve_dst(cx, ins, sizeof(T));
  VM_ROWS(i, g) {  // NormalizeFloatingNumbers.scala:29-38: canonical NaN, -0.0 -> 0.0
    T x = opnd_ld<T>(a, i);
    if (x != x) x = (T)__longlong_as_double(0x7ff8000000000000LL);
    else if (x == (T)0) x = (T)0;
    dst_st<T>(d, i, x, opnd_valid(a, i, g));
  }
}

__device__ __forceinline__ void vm_year(const VMCtx& cx, const VMInstr& ins) {
  Opnd a = resolve(cx, ins.a, 4);
  Dst d = resolve_dst(cx, ins, 4);
  VM_ROWS(i, g) {  // proleptic Gregorian civil-from-days
    int z = opnd_ld<int32_t>(a, i) + 719468;
    int era = (z >= 0 ? z : z - 146096) / 146097;
    unsigned doe = (unsigned)(z - era * 146097);
    unsigned yoe = (doe - doe / 1460 + doe / 36524 - doe / 146096) / 365;
    int y = (int)yoe + era * 400;
    unsigned doy = doe - (365 * yoe + yoe / 4 - yoe / 100);
    unsigned mp = (5 * doy + 2) / 153;
    int m = mp < 10 ? mp + 3 : mp - 9;
    dst_st<int32_t>(d, i, y + (m <= 2), opnd_valid(a, i, g));
  }
}

__device__ __forceinline__ void vm_isnull(const VMCtx& cx, const VMInstr& ins) {
  Opnd a = resolve(cx, ins.a, mt_width(ins.mt));
  Dst d = resolve_dst(cx, ins, 1);
  const bool want_null = ins.op == V_ISNULL;
  VM_ROWS(i, g) { dst_st<int8_t>(d, i, (int8_t)(opnd_valid(a, i, g) != want_null), true); }
}

#define VM_DISPATCH_INT_FLOAT(fn)                      \
  switch (ins.mt) {                                    \
    case MT_I8: fn<int8_t>(cx, ins); break;            \
    case MT_I16: fn<int16_t>(cx, ins); break;          \
    case MT_I32: fn<int32_t>(cx, ins); break;          \
    case MT_I64: fn<int64_t>(cx, ins); break;          \
    case MT_F32: fn<float>(cx, ins); break;            \
    case MT_F64: fn<double>(cx, ins); break;           \
    default: break;                                    \
  }
#define VM_DISPATCH_ALL(fn)                            \
  switch (ins.mt) {                                    \
    case MT_I8: fn<int8_t>(cx, ins); break;            \
    case MT_I16: fn<int16_t>(cx, ins); break;          \
    case MT_I32: fn<int32_t>(cx, ins); break;          \
    case MT_I64: fn<int64_t>(cx, ins); break;          \
    case MT_I128: fn<i128>(cx, ins); break;            \
    case MT_F32: fn<float>(cx, ins); break;            \
    default: fn<double>(cx, ins); break;               \
  }

// Run the whole program over this CTA's tile.  No barrier is needed: registers are thread private.
static __device__ __noinline__ void vm_run(const VMCtx& cx, const VMInstr* __restrict__ code) {
  const int n = cx.hdr->ninstr;
  for (int pc = 0; pc < n; pc++) {
    const VMInstr& ins = code[pc];
    switch (ins.op) {
      case V_ADD: case V_SUB: case V_MUL: case V_DIV: case V_MOD: case V_PMOD:
        if (ins.mt == MT_I128) vm_arith128(cx, ins); else { VM_DISPATCH_INT_FLOAT(vm_arith) }
        break;
      case V_EQ: case V_NE: case V_LT: case V_LE: case V_GT: case V_GE: case V_EQNS:
        VM_DISPATCH_ALL(vm_compare)
        break;
      case V_AND: case V_OR: case V_NOT: vm_logic(cx, ins); break;
      case V_ISNULL: case V_ISNOTNULL: vm_isnull(cx, ins); break;
      case V_COALESCE: case V_IF: case V_MOV: case V_NEG: case V_ABS:
        VM_DISPATCH_ALL(vm_select)
        break;
      case V_CAST: VM_DISPATCH_ALL(vm_cast1) break;
      case V_RESCALE_UP: case V_RESCALE_DOWN: case V_CHECK_PREC:
        switch (ins.mt) {
          case MT_I32: vm_decimal<int32_t>(cx, ins); break;
          case MT_I64: vm_decimal<int64_t>(cx, ins); break;
          default: vm_decimal<i128>(cx, ins); break;
        }
        break;
      case V_MULDEC: vm_muldec(cx, ins); break;
      case V_DEC2F64:
        switch (ins.mt) {
          case MT_I32: vm_dec2f64<int32_t>(cx, ins); break;
          case MT_I64: vm_dec2f64<int64_t>(cx, ins); break;
          default: vm_dec2f64<i128>(cx, ins); break;
        }
        break;
      case V_NORM_NAN_ZERO: if (ins.mt == MT_F32) vm_normnz<float>(cx, ins); else vm_normnz<double>(cx, ins); break;
      case V_YEAR: vm_year(cx, ins); break;
      default: break;
    }
  }
}

constexpr int VM_SMEM_CODE = 64;  // instructions cached in shared memory
struct VMShared {
  VMProgramHeader hdr;
  VMInstr code[VM_SMEM_CODE];
};
// cooperative copy of the program into shared memory; returns the code pointer to execute from
static __device__ __forceinline__ const VMInstr* vm_load_program(VMShared& sh, const VMProgramHeader* g_hdr,
                                                                 const VMInstr* g_code) {
  const int* src = reinterpret_cast<const int*>(g_hdr);
  int* dst = reinterpret_cast<int*>(&sh.hdr);
  for (int k = threadIdx.x; k < (int)(sizeof(VMProgramHeader) / 4); k += blockDim.x) dst[k] = src[k];
  __syncthreads();
  const int n = sh.hdr.ninstr;
  if (n > VM_SMEM_CODE) return g_code;
  const int4* s4 = reinterpret_cast<const int4*>(g_code);
  int4* d4 = reinterpret_cast<int4*>(sh.code);
  for (int k = threadIdx.x; k < n * (int)(sizeof(VMInstr) / 16); k += blockDim.x) d4[k] = s4[k];
  __syncthreads();
  return sh.code;
}

// read output `o` of the program for tile row i (global row g)
template <typename T>
__device__ __forceinline__ T vm_out(const VMCtx& cx, const Opnd& o, int i) { return opnd_ld<T>(o, i); }
#endif  // __CUDACC__

// host-side compiled program
struct Program {
  VMProgramHeader hdr;
  std::vector<VMInstr> code;
  std::vector<int> col_dtype;          // expected dtype of each referenced input column (-1 unused)
  std::vector<int> out_dtype, out_scale, out_precision;
  std::vector<uint8_t> out_nullable;
  DevBuf d_hdr, d_code;                // device copies
};
Program* program_from(b2_handle h);

}  // namespace b2
