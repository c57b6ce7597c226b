// TEST INFRASTRUCTURE ONLY: executes the LayerNorm kernels of tvts_b200/csrc/layernorm.cu on the CPU SIMT stand-in (widths 128*NV,
// including NV = 5 (640, added for the H/14 test model) and the fused bias-gradient column sums of the backward, neither of which has
// been through a GPU test yet).
//   g++ -O1 -std=c++20 -pthread -shared -fPIC -DTVTS_HOST_SHIM -I tests/host_kernels harness_ln.cpp
#include "host_simt.h"

namespace ln {
#include "../../tvts_b200/csrc/layernorm.cu"

template <int NV>
void fwd(const float* x, const float* g, const float* b, void* y, int y_bf16, float* mean, float* rstd, long long M, float eps) {
  const unsigned grid = (unsigned)((M + kWarps - 1) / kWarps);
  simt::launch(grid, 1, 1, kWarps * 32, [&] {
    if (y_bf16) ln_fwd_kernel<NV, true>(x, g, b, y, mean, rstd, M, eps);
    else ln_fwd_kernel<NV, false>(x, g, b, y, mean, rstd, M, eps);
  });
}
template <int NV>
void bwd(const void* dy, int dy_bf16, const float* x, const float* mean, const float* rstd, const float* g, const float* r1, const float* r2,
         float* dx, void* dxb, float* dg, float* db, float* dxsum, long long M, unsigned grid) {
  simt::launch(grid, 1, 1, kWarps * 32, [&] {
    if (dy_bf16) ln_bwd_kernel<NV, true>(dy, x, mean, rstd, g, r1, r2, dx, (bf16*)dxb, dg, db, dxsum, M);
    else ln_bwd_kernel<NV, false>(dy, x, mean, rstd, g, r1, r2, dx, (bf16*)dxb, dg, db, dxsum, M);
  });
}
}  // namespace ln

#define LN_SWITCH(NVV, CALL) \
  switch (NVV) { case 1: { constexpr int NV = 1; CALL; } break; case 2: { constexpr int NV = 2; CALL; } break; case 4: { constexpr int NV = 4; CALL; } break; \
                 case 5: { constexpr int NV = 5; CALL; } break; case 6: { constexpr int NV = 6; CALL; } break; case 8: { constexpr int NV = 8; CALL; } break; \
                 case 10: { constexpr int NV = 10; CALL; } break; default: return -1; }

extern "C" {
int h_ln_fwd(const float* x, const float* g, const float* b, void* y, int y_bf16, float* mean, float* rstd, long long M, long long D, float eps) {
  LN_SWITCH((int)(D / 128), ln::fwd<NV>(x, g, b, y, y_bf16, mean, rstd, M, eps));
  return 0;
}
int h_ln_bwd(const void* dy, int dy_bf16, const float* x, const float* mean, const float* rstd, const float* g, const float* r1, const float* r2,
             float* dx, void* dxb, float* dg, float* db, float* dxsum, long long M, long long D, int grid) {
  LN_SWITCH((int)(D / 128), ln::bwd<NV>(dy, dy_bf16, x, mean, rstd, g, r1, r2, dx, dxb, dg, db, dxsum, M, (unsigned)grid));
  return 0;
}
}
