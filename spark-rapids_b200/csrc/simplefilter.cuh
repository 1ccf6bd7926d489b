// simplefilter.cuh — the "simple predicate" shape shared by the selection-vector kernel (simplefilter.cu) and the fused
// filter + probe kernel (join.cu): a conjunction of comparisons between NOT NULL fixed-width integer columns and literals.
#pragma once
#include <cstdint>
namespace b2 {
struct Program;
struct Table;
constexpr int SF_MAX_TERMS = 8;
struct SimpleTerm { const void* col; int32_t width; int32_t truth; int64_t lit; };   // truth: bit0 '<', bit1 '==', bit2 '>'
struct SimplePred { int32_t n; int32_t pad; SimpleTerm t[SF_MAX_TERMS]; };
// pattern match of a compiled predicate program against a batch; false = not of the simple shape (the VM evaluates it)
bool simple_pred_of(const Program* prog, const Table* t, SimplePred& sp);
}  // namespace b2
