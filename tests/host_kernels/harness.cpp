// TEST INFRASTRUCTURE ONLY: executes the streaming kernels of tvts_b200/csrc/{v1_glue,input_stage}.cu on the CPU, thread by thread over
// the same grid the real launchers use, so tests/test_host_kernels_cpu.py can compare their index arithmetic with the torch restatements
// without a GPU.  Built by the test with:  g++ -O1 -shared -fPIC -DTVTS_HOST_SHIM -I tests/host_kernels harness.cpp
#include "host_shim.h"

namespace v1g {
#include "../../tvts_b200/csrc/v1_glue.cu"
}
namespace ins {
#include "../../tvts_b200/csrc/input_stage.cu"
}

template <typename F>
static void launch(unsigned gx, unsigned gy, unsigned block, F body) {
  gridDim.x = gx; gridDim.y = gy; blockDim.x = block;
  for (unsigned by = 0; by < gy; ++by)
    for (unsigned bx = 0; bx < gx; ++bx)
      for (unsigned tx = 0; tx < block; ++tx) {
        blockIdx.x = bx; blockIdx.y = by; threadIdx.x = tx;
        body();
      }
}
static unsigned blocks(long long total, int threads) { return (unsigned)((total + threads - 1) / threads); }

extern "C" {

void h_tubelet_gather(const float* video, const long long* keep, void* cols, long long B, long long T, long long R, long long p, long long n) {
  const long long total = B * (T / 2) * n * 3 * 2 * p * (p / 4);
  launch(blocks(total, 256), 1, 256, [&] { v1g::tubelet_gather_kernel(video, keep, (bf16*)cols, (int)T, (int)R, (int)p, (int)n, total); });
}
void h_assemble_tube(const float* tok, const float* cls, const float* pos, const float* tem, const long long* keep, float* x0, long long B,
                     long long nt, long long n, long long D) {
  const long long total = B * (1 + nt * n) * (D / 4);
  launch(blocks(total, 256), 1, 256, [&] { v1g::assemble_tube_kernel(tok, cls, pos, tem, keep, x0, (int)nt, (int)n, (int)(D / 4), total); });
}
void h_assemble_tube_bwd(const float* dx0, const long long* keep, float* dcls, float* dpos, float* dtem, void* dtok, long long B, long long nt,
                         long long n, long long D) {
  launch((unsigned)(nt + 1), (unsigned)B, 192, [&] { v1g::assemble_tube_bwd_kernel(dx0, keep, dcls, dpos, dtem, (bf16*)dtok, (int)nt, (int)n, (int)(D / 4)); });
}
void h_relu_bf16(const float* x, void* y, long long n) {
  launch(blocks(n / 4, 256), 1, 256, [&] { v1g::relu_bf16_kernel(x, (bf16*)y, n / 4); });
}
void h_relu_bwd(const float* x, const float* dy, float* dx, long long n) {
  launch(blocks(n / 4, 256), 1, 256, [&] { v1g::relu_bwd_kernel(x, dy, dx, n / 4); });
}
void h_patch_gather_u8(const unsigned char* video, const long long* keep, void* cols, long long B, long long T, long long R, long long p,
                       long long n, const float* mean3, const float* std3) {
  const long long total = B * T * n * 3 * p * (p / 4);
  launch(blocks(total, 256), 1, 256, [&] {
    ins::patch_gather_u8_kernel(video, keep, (bf16*)cols, (int)T, (int)R, (int)p, (int)n, mean3[0], mean3[1], mean3[2], std3[0], std3[1],
                                std3[2], total);
  });
}
void h_patch_gather_ld(const float* video, const long long* keep, void* cols, long long B, long long T, long long R, long long p, long long n,
                       long long ld) {
  const long long total = B * T * n * (ld / 2);
  launch(blocks(total, 256), 1, 256, [&] { ins::patch_gather_ld_kernel(video, keep, (bf16*)cols, (int)B, (int)T, (int)R, (int)p, (int)n, (int)ld, total); });
}
void h_cast_pad(const float* src, void* dst, long long rows, long long cols, long long ld) {
  const long long total = rows * (ld / 2);
  launch(blocks(total, 256), 1, 256, [&] { ins::cast_pad_kernel(src, (bf16*)dst, (int)cols, (int)ld, total); });
}

}  // extern "C"
