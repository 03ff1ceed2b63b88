// CPU functional model of the primitive policy (the counterpart of piper_b200/csrc/tc_policy_dev.cuh) - test
// infrastructure.  See tests/sim/mrf_sim.cpp for what the model covers and what it cannot.
#pragma once
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <functional>
#include <limits>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "../../piper_b200/csrc/kernels.cuh"   // TmapDesc

namespace simtc {

struct SimAbort : std::runtime_error { using std::runtime_error::runtime_error; };

struct SimCta;
struct SimMbar {
  std::mutex m;
  std::condition_variable cv;
  int count = 0, pending = 0;
  long long tx = 0;
  uint32_t phase = 0;
  const char* name = "?";
};

struct CtaBarrier {   // __syncthreads for n threads, reusable
  std::mutex m; std::condition_variable cv; int n = 0, waiting = 0; unsigned gen = 0;
  void arrive_and_wait(std::atomic<bool>& abort) {
    std::unique_lock<std::mutex> l(m);
    const unsigned g = gen;
    if (++waiting == n) { waiting = 0; ++gen; cv.notify_all(); return; }
    while (gen == g) {
      if (cv.wait_for(l, std::chrono::seconds(60)) == std::cv_status::timeout || abort.load())
        if (gen == g) { abort = true; throw SimAbort("__syncthreads never completed"); }
    }
  }
};

// An in-order asynchronous engine (the tensor core, the TMA unit): operations are queued by the issuing thread and executed
// LATER by a worker, with random pauses, so that (a) results the kernel reads before waiting on the right mbarrier are
// stale and (b) operands the kernel overwrites before the engine has consumed them are seen overwritten - the two ways a
// pipelined kernel's barrier protocol can be insufficient even though it never deadlocks.
struct AsyncEngine {
  std::mutex m;
  std::condition_variable cv;
  std::vector<std::function<void()>> q;
  size_t head = 0;
  bool stop = false;
  uint32_t rng = 12345u;
  std::thread worker;
  void start(uint32_t seed) {
    rng = seed * 2654435761u + 1u;
    worker = std::thread([this] {
      for (;;) {
        std::function<void()> op;
        {
          std::unique_lock<std::mutex> l(m);
          cv.wait(l, [this] { return stop || head < q.size(); });
          if (head >= q.size()) return;                 // stop requested and the queue is drained
          op = std::move(q[head++]);
        }
        rng = rng * 1664525u + 1013904223u;
        if ((rng >> 28) == 0) std::this_thread::sleep_for(std::chrono::microseconds(50 + ((rng >> 16) & 255)));
        op();
      }
    });
  }
  void push(std::function<void()> op) {
    { std::lock_guard<std::mutex> l(m); q.push_back(std::move(op)); }
    cv.notify_one();
  }
  void finish() {
    { std::lock_guard<std::mutex> l(m); stop = true; }
    cv.notify_one();
    if (worker.joinable()) worker.join();
  }
};

struct SimCta {
  uint8_t* smem = nullptr;            // 128-byte aligned, smem_bytes long
  int smem_bytes = 0;
  int control_warps = 3;              // warps that run converged on hardware (TMA x 2, MMA): only lane 0 blocks in the model
  float tmem[128][512];
  uint32_t tmem_base = 0xdeadbeef;
  uint32_t tmem_cols = 0;
  CtaBarrier sync;
  AsyncEngine tensor_core, tma;       // in-order, asynchronous (see AsyncEngine)
  CtaBarrier named[16];               // bar.sync id, count (count set on first use)
  int aligned_ops[1024] = {};         // per thread: .sync.aligned collectives executed (tcgen05.ld / st) - must agree per warp
  std::atomic<bool> abort{false};
  std::mutex err_m;
  std::string err;
  void fail(const std::string& e) {
    std::lock_guard<std::mutex> l(err_m);
    if (err.empty()) err = e;
    abort = true;
  }
};

inline uint16_t f2bf(float f) {
  uint32_t u; memcpy(&u, &f, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return uint16_t((u >> 16) | 0x40);
  return uint16_t((u + 0x7fffu + ((u >> 16) & 1u)) >> 16);
}
inline float bf2f(uint16_t h) { uint32_t u = uint32_t(h) << 16; float f; memcpy(&f, &u, 4); return f; }
inline float h2f(uint16_t h) {                                // IEEE binary16 -> float
  const uint32_t sign = uint32_t(h & 0x8000u) << 16, e = (h >> 10) & 0x1fu, m = h & 0x3ffu;
  uint32_t x;
  if (e == 0) {
    if (m == 0) x = sign;
    else { int sh = 0; uint32_t mm = m; while (!(mm & 0x400u)) { mm <<= 1; ++sh; } x = sign | (uint32_t(113 - sh) << 23) | ((mm & 0x3ffu) << 13); }
  } else if (e == 31) x = sign | 0x7f800000u | (m << 13);
  else x = sign | ((e + 112) << 23) | (m << 13);
  float f; memcpy(&f, &x, 4); return f;
}
inline uint16_t f2h(float f) {                                // round to nearest even, overflow -> inf
  uint32_t x; memcpy(&x, &f, 4);
  const uint32_t sign = (x >> 16) & 0x8000u;
  x &= 0x7fffffffu;
  if (x >= 0x7f800000u) return uint16_t(sign | 0x7c00u | (x > 0x7f800000u ? 0x200u : 0u));
  if (x >= 0x477ff000u) return uint16_t(sign | 0x7c00u);
  if (x < 0x33000001u) return uint16_t(sign);
  const int e = int(x >> 23) - 127;
  uint32_t m = (x & 0x7fffffu) | 0x800000u, base; int shift;
  if (e < -14) { shift = 13 + (-14 - e); base = 0; } else { shift = 13; base = uint32_t(e + 15) << 10; m &= 0x7fffffu; }
  const uint32_t q = m >> shift, rem = m & ((1u << shift) - 1), halfway = 1u << (shift - 1);
  uint32_t r = base + q;
  if (rem > halfway || (rem == halfway && (q & 1u))) ++r;
  return uint16_t(sign | r);
}

struct SimPrim {
  static constexpr bool kSim = true;
  using Mbar = SimMbar;
  struct Ctx {
    int tid_, block_, grid_;
    SimCta* cta;
    int tid() const { return tid_; }
    int block() const { return block_; }
    int grid() const { return grid_; }
  };
  static void check(Ctx& c, bool ok, const char* what) {
    if (!ok) { c.cta->fail(std::string("check failed: ") + what + " (tid " + std::to_string(c.tid_) + ")"); throw SimAbort(what); }
  }
  static uint32_t saddr(Ctx& c, const void* p) {
    const long long off = (const uint8_t*)p - c.cta->smem;
    check(c, off >= 0 && off < c.cta->smem_bytes, "pointer outside shared memory");
    return (uint32_t)off;
  }
  static int bcast0(Ctx&, int v) { return v; }           // values broadcast in the kernel are warp-uniform by construction
  static void syncwarp() {}
  static void syncthreads(Ctx& c) { c.cta->sync.arrive_and_wait(c.cta->abort); }
  static void bar_sync(Ctx& c, int id, int count) {
    check(c, id > 0 && id < 16 && count % 32 == 0, "named barrier id / count");
    CtaBarrier& b = c.cta->named[id];
    { std::lock_guard<std::mutex> l(b.m); if (b.n == 0) b.n = count; }
    check(c, b.n == count, "named barrier used with two different thread counts");
    b.arrive_and_wait(c.cta->abort);
  }
  static bool elect_one(Ctx& c) { return (c.tid_ & 31) == 0; }
  static long long clock() { return 0; }
  static void prof_add(unsigned long long*, long long) {}
  static void pdl_launch() {}
  static void pdl_sync() {}
  static void fence_mbar_init() {}
  static void fence_async_proxy() {}
  static void fence_tc_before() {}
  static void fence_tc_after() {}

  static void mbar_init(Ctx&, Mbar* m, uint32_t count) { m->count = m->pending = (int)count; m->tx = 0; m->phase = 0; }
  static void complete_locked(Mbar* m) {
    if (m->pending == 0 && m->tx == 0) { m->phase ^= 1u; m->pending = m->count; m->cv.notify_all(); }
  }
  static void mbar_arrive(Ctx& c, Mbar* m) {
    std::lock_guard<std::mutex> l(m->m);
    check(c, m->pending > 0, "more arrivals than the barrier expects in one phase");
    --m->pending;
    complete_locked(m);
  }
  static void mbar_expect_tx(Ctx& c, Mbar* m, uint32_t bytes) {
    std::lock_guard<std::mutex> l(m->m);
    check(c, m->pending > 0, "more arrivals than the barrier expects in one phase");
    m->tx += bytes;
    --m->pending;
    complete_locked(m);
  }
  static void complete_tx(Mbar* m, uint32_t bytes) {
    std::lock_guard<std::mutex> l(m->m);
    m->tx -= bytes;
    complete_locked(m);
  }
  static void mbar_wait(Ctx& c, Mbar* m, uint32_t parity) {
    // Control warps (TMA x 2, MMA) run converged on hardware: all 32 lanes observe a barrier phase together and only the
    // elected lane acts.  std::threads are not in lockstep - a lane that reaches a parity wait late could find the phase
    // flipped twice and wait forever - so in the model only the elected lane of a control warp blocks.
    if ((c.tid_ >> 5) < c.cta->control_warps && (c.tid_ & 31) != 0) return;
    std::unique_lock<std::mutex> l(m->m);
    const auto deadline = std::chrono::steady_clock::now() + std::chrono::seconds(60);
    while (m->phase == parity) {                          // the phase with this parity has not completed yet
      if (c.cta->abort.load()) throw SimAbort("aborted");
      if (m->cv.wait_until(l, std::chrono::steady_clock::now() + std::chrono::milliseconds(200)) == std::cv_status::timeout &&
          std::chrono::steady_clock::now() > deadline) {
        l.unlock();
        c.cta->fail("deadlock: thread " + std::to_string(c.tid_) + " (warp " + std::to_string(c.tid_ >> 5) +
                    ") waited 60 s on an mbarrier, parity " + std::to_string(parity));
        throw SimAbort("deadlock");
      }
    }
  }
  static void bulk_g2s(Ctx& c, uint32_t dst, const void* src, uint32_t bytes, Mbar* m) {
    check(c, (dst & 15) == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0 && (bytes & 15) == 0 && bytes > 0,
          "cp.async.bulk needs 16-byte aligned addresses and size");
    check(c, dst + bytes <= (uint32_t)c.cta->smem_bytes, "cp.async.bulk writes past shared memory");
    SimCta* cta = c.cta;
    cta->tma.push([cta, dst, src, bytes, m] {            // lands some time later; the mbarrier is how the kernel finds out
      memcpy(cta->smem + dst, src, bytes);
      complete_tx(m, bytes);
    });
  }
  // cp.async.bulk.tensor.3d (tiled): a dense box, elements outside the tensor read as zero, full box bytes credited
  using TensorMap = pb200::TmapDesc;
  static void tma_load_3d(Ctx& c, uint32_t dst, const TensorMap* tm, int x, int y, int z, Mbar* m) {
    const TensorMap t = *tm;
    const uint32_t bytes = uint32_t(t.box[0]) * t.box[1] * t.box[2] * 4u;
    check(c, (dst & 127) == 0, "cp.async.bulk.tensor needs a 128-byte aligned shared-memory destination");
    check(c, ((x * 4) & 15) == 0, "cp.async.bulk.tensor: innermost start coordinate must be 16-byte aligned (illegal instruction on hardware)");
    check(c, (reinterpret_cast<uintptr_t>(t.base) & 15) == 0 && (t.stride1 & 15) == 0 && (t.stride2 & 15) == 0,
          "tensor map: base and strides must be 16-byte aligned");
    check(c, t.box[0] > 0 && t.box[0] <= 256 && t.box[1] > 0 && t.box[1] <= 256 && (t.box[0] * 4) % 16 == 0, "tensor map: box limits");
    check(c, dst + bytes <= (uint32_t)c.cta->smem_bytes, "cp.async.bulk.tensor writes past shared memory");
    SimCta* cta = c.cta;
    cta->tma.push([cta, dst, t, x, y, z, bytes, m] {
      float* d = reinterpret_cast<float*>(cta->smem + dst);
      for (int k = 0; k < t.box[2]; ++k)
        for (int j = 0; j < t.box[1]; ++j)
          for (int i = 0; i < t.box[0]; ++i) {
            const int xi = x + i, yj = y + j, zk = z + k;
            const bool in = xi >= 0 && xi < t.dims[0] && yj >= 0 && yj < t.dims[1] && zk >= 0 && zk < t.dims[2];
            *d++ = in ? *reinterpret_cast<const float*>(reinterpret_cast<const uint8_t*>(t.base) + (long long)zk * t.stride2 +
                                                        (long long)yj * t.stride1 + (long long)xi * 4)
                      : 0.f;
          }
      complete_tx(m, bytes);
    });
  }
  static void tma_prefetch_desc(const TensorMap*) {}
  static void tmem_alloc(Ctx& c, uint32_t* slot, uint32_t cols) {
    check(c, cols >= 32 && cols <= 512 && (cols & (cols - 1)) == 0, "tmem columns must be a power of two in [32, 512]");
    c.cta->tmem_cols = cols;
    *slot = 0;
  }
  static void tmem_dealloc(Ctx&, uint32_t, uint32_t) {}
  // D[128][N] (+)= A[128][K] * B[N][K]^T with K = 32 bytes of elements; operands K-major, no swizzle:
  //   element (row, k) at  start + (k / E) * LBO + (row / 8) * 128 + (row % 8) * 16 + (k % E) * elem_bytes,  E = 16 / elem_bytes
  template <bool TF32>
  static void mma_any(Ctx& c, uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t acc) {
    constexpr int ES = TF32 ? 4 : 2, E = 16 / ES, K = 2 * E;
    const uint32_t a0 = (a_lo & 0x3FFFu) << 4, albo = ((a_lo >> 16) & 0x3FFFu) << 4;
    const uint32_t b0 = (b_lo & 0x3FFFu) << 4, blbo = ((b_lo >> 16) & 0x3FFFu) << 4;
    const int N = int((idesc >> 17) & 0x3Fu) << 3, M = int((idesc >> 24) & 0x1Fu) << 4;
    check(c, M == 128 && N >= 16 && N <= 256 && N % 16 == 0, "instruction descriptor shape");
    const uint32_t fmt = (idesc >> 7) & 7u;               // kind::f16: 0 = F16, 1 = BF16; kind::tf32: 2
    check(c, ((idesc >> 4) & 3u) == 1 && ((idesc >> 10) & 7u) == fmt && (TF32 ? fmt == 2 : fmt <= 1), "instruction descriptor formats");
    const uint32_t col0 = tmem_d & 0xFFFFu;
    check(c, (tmem_d >> 16) == 0 && col0 + N <= c.cta->tmem_cols, "accumulator address outside the TMEM allocation");
    auto elem = [&](uint32_t start, uint32_t lbo, int row, int k) -> float {
      const uint32_t addr = start + uint32_t(k / E) * lbo + uint32_t(row / 8) * 128u + uint32_t(row % 8) * 16u + uint32_t(k % E) * ES;
      if (addr + ES > (uint32_t)c.cta->smem_bytes) { c.cta->fail("tcgen05.mma operand outside shared memory"); throw SimAbort("operand"); }
      if (TF32) {
        uint32_t u; memcpy(&u, c.cta->smem + addr, 4);
        u &= 0xFFFFE000u;                                  // the tensor core reads 19 bits of a tf32 operand
        float f; memcpy(&f, &u, 4);
        return f;
      }
      uint16_t h; memcpy(&h, c.cta->smem + addr, 2);
      return fmt == 0 ? h2f(h) : bf2f(h);
    };
    std::vector<float> B(size_t(N) * K);
    for (int n = 0; n < N; ++n) for (int k = 0; k < K; ++k) B[size_t(n) * K + k] = elem(b0, blbo, n, k);
    for (int m = 0; m < 128; ++m) {
      float A[16];
      for (int k = 0; k < K; ++k) A[k] = elem(a0, albo, m, k);
      for (int n = 0; n < N; ++n) {
        float s = 0.f;
        for (int k = 0; k < K; ++k) s += A[k] * B[size_t(n) * K + k];
        float& d = c.cta->tmem[m][col0 + n];
        d = acc ? d + s : s;
      }
    }
  }
  // issue = enqueue: the operands are read and the accumulator written when the tensor core gets to it
  static void mma_f16(Ctx& c, uint32_t d, uint32_t a, uint32_t b, uint32_t idesc, uint32_t acc) {
    Ctx cc = c;
    c.cta->tensor_core.push([cc, d, a, b, idesc, acc]() mutable { try { mma_any<false>(cc, d, a, b, idesc, acc); } catch (const SimAbort&) {} });
  }
  static void mma_tf32(Ctx& c, uint32_t d, uint32_t a, uint32_t b, uint32_t idesc, uint32_t acc) {
    Ctx cc = c;
    c.cta->tensor_core.push([cc, d, a, b, idesc, acc]() mutable { try { mma_any<true>(cc, d, a, b, idesc, acc); } catch (const SimAbort&) {} });
  }
  static float to_tf32(float v) {                          // cvt.rna.tf32.f32: nearest, ties away from zero
    uint32_t u; memcpy(&u, &v, 4);
    if ((u & 0x7f800000u) == 0x7f800000u) return v;
    u = (u + 0x1000u) & 0xffffe000u;
    float f; memcpy(&f, &u, 4);
    return f;
  }
  static float ldg(const float* p) { return *p; }
  // tcgen05.commit: the arrival happens when every MMA issued before it has completed (the engine is in order)
  static void mma_commit(Ctx& c, Mbar* m) {
    Ctx cc = c;
    c.cta->tensor_core.push([cc, m]() mutable { try { mbar_arrive(cc, m); } catch (const SimAbort&) {} });
  }
  static void tmem_ld16(Ctx& c, uint32_t taddr, float (&v)[16]) {
    const uint32_t lane0 = taddr >> 16, col = taddr & 0xFFFFu;
    check(c, lane0 == 32u * ((c.tid_ >> 5) & 3), "tcgen05.ld outside the warp's TMEM lane quadrant (warp id % 4)");
    check(c, col + 16 <= c.cta->tmem_cols, "tcgen05.ld column range");
    ++c.cta->aligned_ops[c.tid_];
    for (int i = 0; i < 16; ++i) v[i] = c.cta->tmem[lane0 + (c.tid_ & 31)][col + i];
  }
  static void tmem_ld16x2(Ctx& c, uint32_t taddr0, uint32_t taddr1, float (&v)[16], float (&w)[16]) {
    tmem_ld16(c, taddr0, v);
    tmem_ld16(c, taddr1, w);
  }
  static void tmem_st16(Ctx& c, uint32_t taddr, const float* v) {
    const uint32_t lane0 = taddr >> 16, col = taddr & 0xFFFFu;
    check(c, lane0 == 32u * ((c.tid_ >> 5) & 3), "tcgen05.st outside the warp's TMEM lane quadrant (warp id % 4)");
    check(c, col + 16 <= c.cta->tmem_cols, "tcgen05.st column range");
    ++c.cta->aligned_ops[c.tid_];
    for (int i = 0; i < 16; ++i) c.cta->tmem[lane0 + (c.tid_ & 31)][col + i] = v[i];
  }
  static void tmem_wait_st() {}
  static void split2_f16(float a, float b, uint32_t& hi, uint32_t& lo) {
    const uint16_t ha = f2h(a), hb = f2h(b);
    hi = uint32_t(ha) | (uint32_t(hb) << 16);
    lo = uint32_t(f2h(a - h2f(ha))) | (uint32_t(f2h(b - h2f(hb))) << 16);
  }
  static void split2_bf16(float a, float b, uint32_t& hi, uint32_t& lo) {
    const uint16_t ha = f2bf(a), hb = f2bf(b);
    hi = uint32_t(ha) | (uint32_t(hb) << 16);
    lo = uint32_t(f2bf(a - bf2f(ha))) | (uint32_t(f2bf(b - bf2f(hb))) << 16);
  }
  static float f16_round(float v) { return h2f(f2h(v)); }
  static uint32_t pack_f16(float a, float b) { return uint32_t(f2h(a)) | (uint32_t(f2h(b)) << 16); }
  static float bf16_round(float v) { return bf2f(f2bf(v)); }
  static uint32_t pack_bf16(float a, float b) { return uint32_t(f2bf(a)) | (uint32_t(f2bf(b)) << 16); }
};


// Run `threads` std::threads of one CTA over `body(ctx)`; returns the first error ("" if none).
template <class Body>
inline std::string run_cta(SimCta& cta, int threads, int block, int grid, Body body) {
  cta.sync.n = threads;
  cta.tensor_core.start(2u * (uint32_t)block + 1u);
  cta.tma.start(2u * (uint32_t)block + 2u);
  std::vector<std::thread> th;
  for (int t = 0; t < threads; ++t)
    th.emplace_back([&, t] {
      SimPrim::Ctx cx{t, block, grid, &cta};
      try {
        body(cx);
      } catch (const SimAbort&) {
      } catch (const std::exception& e) {
        cta.fail(e.what());
      }
    });
  for (auto& t : th) t.join();
  cta.tensor_core.finish();
  cta.tma.finish();
  // tcgen05.ld / st are .sync.aligned: every lane of a warp must execute the same sequence (a lane that skips one hangs the
  // warp on hardware).  Lanes run as independent threads here, so compare their counts afterwards.
  if (cta.err.empty())
    for (int w = 0; w * 32 < threads; ++w)
      for (int l = 1; l < 32 && w * 32 + l < threads; ++l)
        if (cta.aligned_ops[w * 32 + l] != cta.aligned_ops[w * 32]) {
          cta.err = "warp " + std::to_string(w) + ": lanes executed different numbers of tcgen05.ld/st (.sync.aligned): lane 0 " +
                    std::to_string(cta.aligned_ops[w * 32]) + ", lane " + std::to_string(l) + " " + std::to_string(cta.aligned_ops[w * 32 + l]);
          return cta.err;
        }
  return cta.err;
}

struct SmemBuf {                       // 128-byte aligned shared-memory image filled with stale garbage
  std::vector<uint8_t> raw;
  uint8_t* p = nullptr;
  explicit SmemBuf(size_t bytes) : raw(bytes + 128, 0xCD) { p = raw.data() + ((128 - reinterpret_cast<uintptr_t>(raw.data()) % 128) % 128); }
};

}  // namespace simtc
