#pragma once
#include <vector>

#include "ir.h"
#include "vm.h"

namespace b200q {

struct OutDesc { DType type; bool nullable; int slots; };

struct CompiledProgram {
  VmProgram prog;
  std::vector<int> used_cols;   // program column slot -> index in the stage's input schema
  std::vector<OutDesc> outs;
};

// filters: conjuncts in evaluation order; outs: projections (FilterExec/ProjectExec kernel) or
// grouping keys followed by aggregate arguments (HashAgg kernel).
CompiledProgram compile_program(const std::vector<ExprP>& filters, const std::vector<ExprP>& outs, bool with_compact);

// replace column references by the expressions that define them (fusing Project/Filter chains)
ExprP substitute(const ExprP& e, const std::vector<ExprP>& cols);

PhysKind phys_of(const DType& t);

}  // namespace b200q
