// tcgen05 implicit-GEMM convolution (placeholder until the kernel lands: nothing is eligible yet,
// so every conv takes the fp32 SIMT path).
#include "ctx.cuh"

namespace cs {

bool conv_tc_supported(const ConvW&, const Act&) { return false; }

Opd conv_tc_alloc_operand(Arena&, const ConvW&, const Act&) { throw Error(CS_ERR_INVALID, "tcgen05 conv not built"); }

void conv_tc(const Launcher&, const Opd&, const ConvW&, const ConvGeom&, const Epilogue&, Act) {
  throw Error(CS_ERR_INVALID, "tcgen05 conv not built");
}

void pack_tc(cs_ctx*, ConvW&, cudaStream_t) {}

}  // namespace cs
