// Developer probe: which forms of cp.async.bulk.tensor (UTMALDG) does this driver / part accept?  One variant per
// process (a fault kills the context).  Build: nvcc -gencode arch=compute_100a,code=sm_100a -o tma_probe tma_probe.cu
//   tma_probe <variant>   variant = rank(2|3) * 100 + where(0 param, 1 global) * 10 + form(0 plain, 1 .tile + L2 hint)
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s\n", cudaGetErrorString(e), #x); return 2; } } while (0)

__device__ __forceinline__ uint32_t sa(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int RANK, int FORM>
__device__ __forceinline__ void tma(uint32_t dst, const CUtensorMap* tm, int x, int y, int z, uint32_t mbar) {
  if (RANK == 3) {
    if (FORM == 0)
      asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];\n" ::"r"(dst), "l"(tm), "r"(x), "r"(y), "r"(z), "r"(mbar) : "memory");
    else
      asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3, %4}], [%5], %6;\n" ::"r"(dst), "l"(tm), "r"(x), "r"(y), "r"(z), "r"(mbar), "l"(0x1000000000000000ull) : "memory");
  } else {
    if (FORM == 0)
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n" ::"r"(dst), "l"(tm), "r"(x), "r"(y), "r"(mbar) : "memory");
    else
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3}], [%4], %5;\n" ::"r"(dst), "l"(tm), "r"(x), "r"(y), "r"(mbar), "l"(0x1000000000000000ull) : "memory");
  }
}

template <int RANK, int FORM>
__global__ void probe(const __grid_constant__ CUtensorMap tm_param, const CUtensorMap* tm_global, int use_global, int x0, int y0, int z0,
                      int box_floats, float* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  const CUtensorMap* tm = use_global ? tm_global : &tm_param;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(sa(&bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(sa(&bar)), "r"((uint32_t)box_floats * 4u) : "memory");
    tma<RANK, FORM>(sa(smem), tm, x0, y0, z0, sa(&bar));
  }
  // bounded wait
  long long t0 = clock64();
  uint32_t done = 0;
  while (!done) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(done) : "r"(sa(&bar)) : "memory");
    if (clock64() - t0 > 2000000000LL) { if (threadIdx.x == 0) out[0] = -12345.f; return; }
  }
  for (int i = threadIdx.x; i < box_floats; i += blockDim.x) out[i] = reinterpret_cast<float*>(smem)[i];
}

int main(int argc, char** argv) {
  const int variant = argc > 1 ? atoi(argv[1]) : 300;
  const int rank = variant / 100, where = (variant / 10) % 10, form = variant % 10;
  const int box_w = argc > 2 ? atoi(argv[2]) : 132, box_h = argc > 3 ? atoi(argv[3]) : 32, x0 = argc > 4 ? atoi(argv[4]) : -5;
  const int cs = 300, C = 64, B = 2;
  std::vector<float> h(size_t(B) * C * cs);
  for (size_t i = 0; i < h.size(); ++i) h[i] = float(i % 1000) + 0.5f;
  float *d = nullptr, *out = nullptr;
  CK(cudaMalloc(&d, h.size() * 4));
  CK(cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMalloc(&out, 256 * 256 * 4));
  void* fnp = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q));
  using Fn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                          const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  Fn fn = reinterpret_cast<Fn>(fnp);
  CUtensorMap tm;
  CUresult r;
  if (rank == 3) {
    cuuint64_t dims[3] = {cs, C, B}; cuuint64_t strides[2] = {cs * 4, (cuuint64_t)C * cs * 4};
    cuuint32_t box[3] = {(cuuint32_t)box_w, (cuuint32_t)box_h, 1}; cuuint32_t es[3] = {1, 1, 1};
    r = fn(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
           CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  } else {
    cuuint64_t dims[2] = {cs, (cuuint64_t)C * B}; cuuint64_t strides[1] = {cs * 4};
    cuuint32_t box[2] = {(cuuint32_t)box_w, (cuuint32_t)box_h}; cuuint32_t es[2] = {1, 1};
    r = fn(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
           CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  }
  if (r != CUDA_SUCCESS) { printf("variant %d: encode failed %d\n", variant, (int)r); return 3; }
  CUtensorMap* tm_g = nullptr;
  CK(cudaMalloc(&tm_g, sizeof(CUtensorMap)));
  CK(cudaMemcpy(tm_g, &tm, sizeof(CUtensorMap), cudaMemcpyHostToDevice));
  const int n = box_w * box_h;
  const int smem = n * 4 + 1024;
  const int y0 = 16, z0 = 1;
  auto launch = [&](auto kern) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    kern<<<1, 128, smem>>>(tm, tm_g, where, x0, rank == 3 ? y0 : z0 * C + y0, z0, n, out);
  };
  if (rank == 3 && form == 0) launch(probe<3, 0>);
  else if (rank == 3) launch(probe<3, 1>);
  else if (form == 0) launch(probe<2, 0>);
  else launch(probe<2, 1>);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("variant %d box %dx%d x0 %d: FAULT %s\n", variant, box_w, box_h, x0, cudaGetErrorString(e)); return 1; }
  std::vector<float> o(n);
  CK(cudaMemcpy(o.data(), out, n * 4, cudaMemcpyDeviceToHost));
  if (o[0] == -12345.f) { printf("variant %d: TIMEOUT (mbarrier never completed)\n", variant); return 1; }
  int bad = 0;
  for (int j = 0; j < box_h; ++j)
    for (int i = 0; i < box_w; ++i) {
      const int xi = x0 + i;
      const float want = (xi >= 0 && xi < cs) ? h[(size_t(z0) * C + y0 + j) * cs + xi] : 0.f;
      if (o[j * box_w + i] != want) ++bad;
    }
  printf("variant %d box %dx%d x0 %d: ok, %d mismatches\n", variant, box_w, box_h, x0, bad);
  return bad ? 1 : 0;
}
