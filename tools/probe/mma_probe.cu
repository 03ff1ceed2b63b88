// Developer probe: what paces the tcgen05.mma stream of the conv kernel (conv2_body.inl)?  One CTA per SM issues the
// kernel's exact instruction pattern (SS operands, no-swizzle K-major, fp16, M = 128, tap-shifted A start addresses,
// stacked [W_hi; W_lo] B operand) with nothing else running, and prints cycles per k-step (one k-step = the
// instructions that cover K = 16 for one 128-row half).  The tensor-pipe floor is M*N/256 cycles per instruction
// (B300_MICROARCH.md).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -o mma_probe mma_probe.cu      Run: mma_probe [NT [first pattern [last pattern]]]
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s\n", cudaGetErrorString(e), #x); return 2; } } while (0)

__device__ __forceinline__ uint32_t sa(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t desc_lo(uint32_t saddr, uint32_t lbo_bytes) {
  return ((saddr & 0x3FFFF) >> 4) | (((lbo_bytes >> 4) & 0x3FFF) << 16);
}
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {
  return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);   // D = F32, A = B = F16
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile("{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\telect.sync rx|px, %1;\n\t@px mov.s32 %0, 1;\n\t}\n" : "+r"(pred) : "r"(0xFFFFFFFFu));
  return pred != 0;
}
template <int SW>
__device__ __forceinline__ void mma(uint32_t d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t acc) {
  // SW = 0: no swizzle, SBO = 128 B.  SW = 1: 128-byte swizzle, SBO = 1024 B (layout type 2 at bits 61-63 -> hi word bits 29-31)
  constexpr uint32_t HI = SW ? ((1024u >> 4) | (1u << 14) | (2u << 29)) : ((128u >> 4) | (1u << 14));
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tmov.b64 da, {%1, %5};\n\tmov.b64 db, {%2, %5};\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n\t}\n" ::"r"(d), "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(acc), "r"(HI)
      : "memory");
}
__device__ __forceinline__ void commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(sa(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tWAIT_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}\n" ::"r"(sa(bar)),
      "r"(parity)
      : "memory");
}

struct Args {
  int NT, R, k, dil, kb, tiles, pattern, busy, sync;   // sync: per tap, wait on a (ready) mbarrier, fence, commit to a sink mbarrier - as the kernel does
  //   // busy: other warps hammer shared memory (converter-like traffic)
  unsigned long long* out;
};

// patterns
//  0 production: per 128-row half {hi x [W_hi;W_lo] -> D[0,2NT) ; lo x W_hi -> D[NT,2NT)}, two halves
//  1 three instructions per half on matching regions (PIPER_B200_V2_MMA3)
//  2 as 0, correction into a DISJOINT region D[2NT,3NT) (wrong maths; isolates a destination-overlap hazard)
//  3 only the stacked instruction (isolates the N = NT instruction)
//  4 as 0 without tap shift (every tap starts at row 0: isolates start-address alignment)
//  5 as 0, one 128-row half only
//  6 as 0 in the order hi(h0) hi(h1) lo(h0) lo(h1)
//  7 only N = NT instructions, two per half (lo x W_hi twice; isolates the instruction width)
//  8 as 0 with the 128-byte-swizzle descriptor form (K = 64 per 128-byte row; timing only)
template <int pat>
__global__ void __launch_bounds__(256) probe(const Args a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t done, ready, sink;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0);   // warp-uniform role index
  for (int i = tid; i < 160 * 1024 / 4; i += 256) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(sa(&tmem_base_s)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(sa(&done)) : "memory");
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(sa(&ready)) : "memory");
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 100000;\n" ::"r"(sa(&sink)) : "memory");
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}\n" ::"r"(sa(&ready)) : "memory");   // phase 0 complete for good
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tmem_d = (uint32_t)__shfl_sync(0xffffffffu, (int)tmem_base_s, 0);   // warp-uniform: descriptors stay in uniform registers
  __shared__ volatile int stop;
  if (tid == 0) stop = 0;
  __syncthreads();
  if (warp == 0) {
    const int NT = a.NT, R = a.R;
    const uint32_t a_part = pat == 8 ? (uint32_t)R * 128 : (uint32_t)(a.kb * 2) * R * 16;   // one operand half (hi or lo): [K/8][R][16 B]
    const uint32_t a_lbo = (uint32_t)R * 16, w_lbo = 2u * NT * 16;
    const uint32_t a_step = 2u * R, w_step = 4u * NT;
    const uint32_t idesc2 = make_idesc(128, 2 * NT), idesc1 = make_idesc(128, NT);
    const uint32_t ah_base = desc_lo(sa(smem), a_lbo), al_base = desc_lo(sa(smem) + a_part, a_lbo);
    const uint32_t w_base = desc_lo(sa(smem) + 2 * a_part + 1024 - (2 * a_part) % 1024, w_lbo);
    const long long t0 = clock64();
#pragma unroll 1
    for (int tile = 0; tile < a.tiles; ++tile) {
      const uint32_t d_set = tmem_d + (NT <= 64 ? (uint32_t)((tile & 1) * 256) : 0u);   // two accumulator sets when they fit
      uint32_t acc = 0;
#pragma unroll 1
      for (int j = 0; j < a.k; ++j) {
        const uint32_t row = pat == 4 ? 0u : (uint32_t)(j * a.dil) * (pat == 8 ? 8u : 1u);
        uint32_t ah = ah_base + row, al = al_base + row, wb = w_base;
        if (a.sync) { mbar_wait(&ready, 0); asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }
#pragma unroll 1
        for (int kb = 0; kb < a.kb; ++kb) {
          if (elect_one()) {
            const uint32_t h1 = 128u, d1 = d_set + 2u * NT;
            switch (pat) {
              case 1:
                mma<0>(d_set, ah, wb, idesc1, acc); mma<0>(d_set + NT, ah, wb + NT, idesc1, acc); mma<0>(d_set + NT, al, wb, idesc1, 1u);
                mma<0>(d1, ah + h1, wb, idesc1, acc); mma<0>(d1 + NT, ah + h1, wb + NT, idesc1, acc); mma<0>(d1 + NT, al + h1, wb, idesc1, 1u);
                break;
              case 2:
                mma<0>(d_set, ah, wb, idesc2, acc); mma<0>((d_set ^ 256u) + NT, al, wb, idesc1, acc);   // (the other accumulator set)
                mma<0>(d1, ah + h1, wb, idesc2, acc); mma<0>((d1 ^ 256u) + NT, al + h1, wb, idesc1, acc);
                break;
              case 3:
                mma<0>(d_set, ah, wb, idesc2, acc); mma<0>(d1, ah + h1, wb, idesc2, acc);
                break;
              case 5:
                mma<0>(d_set, ah, wb, idesc2, acc); mma<0>(d_set + NT, al, wb, idesc1, 1u);
                break;
              case 6:
                mma<0>(d_set, ah, wb, idesc2, acc); mma<0>(d1, ah + h1, wb, idesc2, acc);
                mma<0>(d_set + NT, al, wb, idesc1, 1u); mma<0>(d1 + NT, al + h1, wb, idesc1, 1u);
                break;
              case 7:
                mma<0>(d_set, al, wb, idesc1, acc); mma<0>(d_set + NT, al, wb, idesc1, acc);
                mma<0>(d1, al + h1, wb, idesc1, acc); mma<0>(d1 + NT, al + h1, wb, idesc1, acc);
                break;
              case 8:
                mma<1>(d_set, ah, wb, idesc2, acc); mma<1>(d_set + NT, al, wb, idesc1, 1u);
                mma<1>(d1, ah + h1 * 8, wb, idesc2, acc); mma<1>(d1 + NT, al + h1 * 8, wb, idesc1, 1u);
                break;
              default:
                mma<0>(d_set, ah, wb, idesc2, acc); mma<0>(d_set + NT, al, wb, idesc1, 1u);
                mma<0>(d1, ah + h1, wb, idesc2, acc); mma<0>(d1 + NT, al + h1, wb, idesc1, 1u);
            }
          }
          __syncwarp();
          acc = 1u;
          if (pat == 8) { ah += 2; al += 2; wb += 2; }   // 32 bytes along the swizzled 128-byte row
          else { ah += a_step; al += a_step; wb += w_step; }
        }
        if (a.sync) { if (elect_one()) commit(&sink); __syncwarp(); }
      }
    }
    const long long t1 = clock64();
    if (elect_one()) commit(&done);
    __syncwarp();
    mbar_wait(&done, 0);
    const long long t2 = clock64();
    stop = 1;
    if (tid == 0 && blockIdx.x == 0) {
      a.out[0] = (unsigned long long)(t1 - t0);
      a.out[1] = (unsigned long long)(t2 - t0);
    }
  } else if (a.busy && warp >= 4) {
    // converter-like traffic: 128-bit shared loads and stores in a scratch area beyond the operands
    uint4* s = reinterpret_cast<uint4*>(smem + 96 * 1024) + (tid - 128);
    uint4 v = make_uint4(tid, 1, 2, 3);
    while (!stop) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const uint4 r = s[i * 128];
        v.x += r.x; v.y ^= r.y;
        s[i * 128 + 1024] = v;
      }
    }
    if (v.x == 0x12345678u) a.out[2] = v.y;
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_d), "r"(512u) : "memory");
}

int main(int argc, char** argv) {
  const int NT = argc > 1 ? atoi(argv[1]) : 64;
  const int pat_lo = argc > 2 ? atoi(argv[2]) : 0, pat_hi = argc > 3 ? atoi(argv[3]) : 7;   // (8, the swizzled form, on request)
  unsigned long long* out;
  CK(cudaMalloc(&out, 64));
  CK(cudaFuncSetAttribute(probe<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  CK(cudaFuncSetAttribute(probe<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  CK(cudaFuncSetAttribute(probe<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  CK(cudaFuncSetAttribute(probe<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  CK(cudaFuncSetAttribute(probe<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  CK(cudaFuncSetAttribute(probe<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  CK(cudaFuncSetAttribute(probe<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  CK(cudaFuncSetAttribute(probe<7>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  CK(cudaFuncSetAttribute(probe<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  static const char* names[] = {"production", "three-instr", "disjoint-corr", "stacked only", "no tap shift", "one half", "hi hi lo lo", "N=NT only", "swizzle128"};
  printf("NT = %d (stacked instruction N = %d): floor %d + %d = %d cycles per half k-step\n", NT, 2 * NT, NT, NT / 2, NT + NT / 2);
  printf("%-14s %3s %3s %4s %4s %5s | %8s %8s %8s\n", "pattern", "k", "dil", "R", "busy", "grid", "instr", "cyc/instr", "cyc/half-kstep");
  struct Cfg { int k, dil, kb; };
  const Cfg cfgs[] = {{7, 3, 2}, {7, 12, 1}, {3, 1, 2}, {1, 1, 4}};
  for (const Cfg& c : cfgs) {
    for (int pat = pat_lo; pat <= pat_hi; ++pat) {
      for (int busy = 0; busy <= 2; ++busy) {   // 2 = busy 0 with the kernel's per-tap synchronisation
        for (int grid : {1, 148}) {
          if ((busy || grid > 1) && pat != 0 && pat != 1) continue;
          if (pat == 2 && NT > 64) continue;
          Args a;
          a.NT = NT; a.k = c.k; a.dil = c.dil; a.kb = c.kb; a.tiles = 64; a.pattern = pat; a.busy = busy == 1; a.sync = busy == 2; a.out = out;
          a.R = 256 + (c.k - 1) * c.dil; a.R = (a.R + 7) / 8 * 8;
          CK(cudaMemset(out, 0, 64));
          switch (pat) {
            case 0: probe<0><<<grid, 256, 180 * 1024>>>(a); break;
            case 1: probe<1><<<grid, 256, 180 * 1024>>>(a); break;
            case 2: probe<2><<<grid, 256, 180 * 1024>>>(a); break;
            case 3: probe<3><<<grid, 256, 180 * 1024>>>(a); break;
            case 4: probe<4><<<grid, 256, 180 * 1024>>>(a); break;
            case 5: probe<5><<<grid, 256, 180 * 1024>>>(a); break;
            case 6: probe<6><<<grid, 256, 180 * 1024>>>(a); break;
            case 7: probe<7><<<grid, 256, 180 * 1024>>>(a); break;
            default: probe<8><<<grid, 256, 180 * 1024>>>(a);
          }
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("%-14s failed: %s\n", names[pat], cudaGetErrorString(e)); return 3; }
          unsigned long long h[2];
          CK(cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost));
          const int per_step = pat == 1 ? 6 : (pat == 3 || pat == 5) ? 2 : 4;
          const double n_instr = double(a.tiles) * c.k * c.kb * per_step;
          const double halves = double(a.tiles) * c.k * c.kb * (pat == 5 ? 1 : 2);
          printf("%-14s %3d %3d %4d %4d %5d | %8.0f %8.1f %8.1f   (issue loop alone %.1f per instr)\n", names[pat], c.k, c.dil, a.R, busy, grid, n_instr,
                 h[1] / n_instr, h[1] / halves, h[0] / n_instr);
        }
      }
    }
  }
  return 0;
}
