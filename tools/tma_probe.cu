// tma_probe.cu -- developer probe: which TMA tiled-load start coordinates work?
// usage: tma_probe W H BX BY x y      (one load per process: a fault kills the context)
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../sbmc_b200/csrc/common.cuh"
using namespace sbmc;

__global__ void probe(const __grid_constant__ CUtensorMap map, float *out, int n, int x, int y) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) uint64_t bar;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_expect_tx(&bar, n * 4);
    tma_load_4d(smem, &map, &bar, x, y, 0, 0);
  }
  mbar_wait(&bar, 0);
  for (int i = threadIdx.x; i < n; i += blockDim.x) out[i] = reinterpret_cast<float *>(smem)[i];
}

int main(int argc, char **argv) {
  int W = atoi(argv[1]), H = atoi(argv[2]), BX = atoi(argv[3]), BY = atoi(argv[4]);
  int x = atoi(argv[5]), y = atoi(argv[6]);
  std::vector<float> h((size_t)W * H);
  for (size_t i = 0; i < h.size(); ++i) h[i] = (float)(i + 1);
  float *d, *o;
  cudaMalloc(&d, h.size() * 4);
  cudaMalloc(&o, (size_t)BX * BY * 4);
  cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  CUtensorMap map;
  const uint64_t dims[4] = {(uint64_t)W, (uint64_t)H, 1, 1};
  const uint64_t strides[3] = {(uint64_t)W * 4, (uint64_t)W * H * 4, (uint64_t)W * H * 4};
  const uint32_t box[4] = {(uint32_t)BX, (uint32_t)BY, 1, 1};
  if (!encode_tensor_map_f32(&map, d, 4, dims, strides, box)) {
    printf("W=%d BX=%d BY=%d x=%d y=%d ENCODE FAILED %s\n", W, BX, BY, x, y, sbmc_b200_last_error());
    return 1;
  }
  probe<<<1, 128, BX * BY * 4>>>(map, o, BX * BY, x, y);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("W=%d BX=%d BY=%d x=%d y=%d FAULT %s\n", W, BX, BY, x, y, cudaGetErrorString(e));
    return 2;
  }
  std::vector<float> r((size_t)BX * BY);
  cudaMemcpy(r.data(), o, r.size() * 4, cudaMemcpyDeviceToHost);
  int bad = 0;
  for (int j = 0; j < BY; ++j)
    for (int i = 0; i < BX; ++i) {
      int xx = x + i, yy = y + j;
      float want = (xx >= 0 && xx < W && yy >= 0 && yy < H) ? h[(size_t)yy * W + xx] : 0.f;
      if (r[(size_t)j * BX + i] != want) ++bad;
    }
  printf("W=%d BX=%d BY=%d x=%d y=%d %s (%d mismatches)\n", W, BX, BY, x, y, bad ? "WRONG" : "ok", bad);
  return bad ? 3 : 0;
}
