// tests/host_emu/emu.cpp -- TEST INFRASTRUCTURE ONLY.
// Compiles the product's TG_HD device functions, functors and host pipeline (mrs_uav_trajectory_generation_b200/csrc)
// with g++ and runs them on CPU threads, so that bit-parity with the oracle can be checked on machines without a GPU
// (`pytest -m "not gpu"`).  The resulting libtg_emu.so exports the same C ABI as libtg_b200.so but is never loaded by
// the package: the product has no CPU path and fails loudly without CUDA.
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <thread>
#include <limits>
#include <vector>

#define TG_VERSION_STRING "tg_emu (host emulation, tests only)"

#include "../../mrs_uav_trajectory_generation_b200/csrc/tg_kernels.cuh"

struct EmuBackend {
  int nthreads;
  explicit EmuBackend(int) {
    const char* e = std::getenv("TG_EMU_THREADS");
    nthreads = e ? std::atoi(e) : (int)std::thread::hardware_concurrency();
    if (nthreads < 1) nthreads = 1;
  }
  void* dev_alloc(size_t n) {
    void* p = std::malloc(n);
    if (!p) throw std::runtime_error("emu: out of memory");
    return p;
  }
  void dev_free(void* p) { std::free(p); }
  void h2d(void* d, const void* s, size_t n) { std::memcpy(d, s, n); }
  void d2h(void* d, const void* s, size_t n) { std::memcpy(d, s, n); }
  void d2d(void* d, const void* s, size_t n) { std::memcpy(d, s, n); }
  void dev_memset(void* d, int v, size_t n) { std::memset(d, v, n); }
  void sync() {}
  double wait_s = 0.0;  // CudaBackend's host-wait trace (TG_TRACE_HOST) has nothing to measure here
  const bool trace_host = false;
  void d2h_overlapped(void* d, const void* s, size_t n) { std::memcpy(d, s, n); }
  // CudaBackend's page-locked arena: plain heap blocks here, released at the next reset
  std::vector<std::unique_ptr<char[]>> pinned_blocks;
  void pinned_reset() { pinned_blocks.clear(); }
  void* pinned_alloc(size_t bytes) {
    pinned_blocks.emplace_back(new char[bytes ? bytes : 1]);
    return pinned_blocks.back().get();
  }
  void d2h_pinned(void* d, const void* s, size_t n) { std::memcpy(d, s, n); }
  void copy_join() {}
  void bind() {}
  void timer_start() {}
  double timer_stop() { return 0.0; }

  template <class W>
  void parallel(size_t n, const W& work) {
    if (n == 0) return;
    const int nt = (int)std::min<size_t>((size_t)nthreads, (n + 63) / 64);
    if (nt <= 1) {
      for (size_t i = 0; i < n; ++i) work(i);
      return;
    }
    std::atomic<size_t> next(0);
    auto run = [&]() {
      for (;;) {
        const size_t i0 = next.fetch_add(64);
        if (i0 >= n) break;
        const size_t i1 = std::min(n, i0 + 64);
        for (size_t i = i0; i < i1; ++i) work(i);
      }
    };
    std::vector<std::thread> th;
    for (int t = 1; t < nt; ++t) th.emplace_back(run);
    run();
    for (auto& t : th) t.join();
  }
  template <class F>
  void for_each(size_t n, const F& f) { parallel(n, [&](size_t i) { f(i); }); }
  // one "warp" per work item (CudaBackend::for_each_warp): TG_PHASE runs the 32 lanes of every phase one after the other
  template <class F>
  void for_each_warp(size_t n, const F& f) { parallel(n, [&](size_t i) { f(i, 0); }); }
  // work items with F::kScratch doubles of private scratch (shared memory on the device), stride 1 here
  template <class F>
  void for_each_scratch(size_t n, const F& f) {
    parallel(n, [&](size_t i) {
      double scratch[F::kScratch];
      f(i, scratch, 1);
    });
  }
  // Mirrors CudaBackend::extrema_refill (persistent warps with lane refill): item by item here
  template <int Q>
  void extrema_refill(int, size_t n_max, const double* coef, const double* times, double* maxima, const int* work, const int* n_dev, int*) {
    for_each_scratch(n_max, tg::ExtremaRawFn<Q>{coef, times, maxima, work, n_dev});
  }
  void fork(int) {}
  void join(int) {}
  template <class F>
  void for_each_scratch_on(int, size_t n, const F& f) { for_each_scratch(n, f); }
  // one "warp" per group of four instances (octet kernel) or per instance (general kernel): phases run lane by lane
  // (lanes own disjoint outputs within a phase).  Mirrors k_solve_oct / k_solve of cuda_backend.cu.
  // same classes as CudaBackend::solve_class on a B200 (228 KB shared memory per SM, 227 KB per CTA, 12 warps by registers)
  int solve_class(int ws_doubles, int oct_ws_doubles) const {
    const size_t per_sm = 233472, optin = 232448;
    if (oct_ws_doubles > 0 && !std::getenv("TG_EMU_NO_OCTET")) {
      const size_t smem = (size_t)4 * oct_ws_doubles * sizeof(double);
      const int w = smem > optin ? 0 : std::min((int)(per_sm / (smem + 1024)), 12);
      if (w >= 1) return w;
    }
    return ((size_t)ws_doubles * sizeof(double) * 4 <= optin) ? 0 : -1;
  }
  bool skip_thread_eligible = false;
  // Mirrors CudaBackend::solve_thread: one work item per instance, slab rows contiguous (element stride 1)
  template <class D>
  void solve_thread(size_t inst_begin, size_t inst_end, int rows_cap, const D& desc) {
    if (inst_end <= inst_begin) return;
    parallel(inst_end - inst_begin, [&](size_t k) {
      tg::SolveInst I;
      if (!desc.instance(inst_begin + k, I) || !tg::thread_eligible(I)) return;
      const double nan = std::numeric_limits<double>::quiet_NaN();
      std::vector<double> slab((size_t)rows_cap * tg::kThrRow, nan);
      tg::solve_thread(I, slab.data(), 1);
    });
  }
  // Mirrors CudaBackend::solve: instances the octet routine can take go through solve_octets in groups of four (one
  // "warp"), the others (when `mixed`) through solve_warp.
  template <class D>
  void solve(size_t inst_begin, size_t inst_end, int ws_doubles, int oct_ws_doubles, int np_cap, bool mixed, const D& desc0) {
    if (inst_end <= inst_begin) return;
    const size_t n_inst = inst_end - inst_begin;
    struct Shifted {  // instance numbering relative to the range
      const D& d;
      size_t off;
      bool instance(size_t i, tg::SolveInst& I) const { return d.instance(i + off, I); }
    } desc{desc0, inst_begin};
    const bool use_oct = solve_class(ws_doubles, oct_ws_doubles) >= 1;
    np_cap = std::max(np_cap, 1);
    auto takes = [&](const tg::SolveInst& I) { return use_oct && tg::octet_eligible(I) && tg::octet_ws_doubles(I.S, I.np) <= oct_ws_doubles && I.np <= np_cap; };
    if (use_oct) {
      if (std::getenv("TG_EMU_TRACE")) std::fprintf(stderr, "[emu] octet solve path: %zu instances\n", n_inst);
      parallel((n_inst + 3) / 4, [&](size_t grp) {
        tg::SolveInst I[4];
        int nmax = 0;
        // NaN-filled: on the device neither the shared-memory ring nor the slab is initialised
        const double nan = std::numeric_limits<double>::quiet_NaN();
        std::vector<double> ws((size_t)4 * oct_ws_doubles, nan), us((size_t)4 * np_cap * tg::kOctURow, nan);
        for (int o = 0; o < 4; ++o) {
          const size_t inst = grp * 4 + o;
          const bool ok = inst < n_inst && desc.instance(inst, I[o]) && !(skip_thread_eligible && tg::thread_eligible(I[o])) && takes(I[o]);
          if (!ok) {
            I[o] = tg::SolveInst{};
            I[o].hbw = tg::kOctHbw;
          }
          tg::octet_ws_bind(I[o], ws.data() + (size_t)o * oct_ws_doubles, us.data() + (size_t)o * np_cap * tg::kOctURow);
          nmax = std::max(nmax, I[o].np);
        }
        if (nmax > 0) tg::solve_octets(I, 0, nmax);
      });
      if (!mixed) return;
    }
    parallel(n_inst, [&](size_t inst) {
      tg::SolveInst I;
      if (!desc.instance(inst, I)) return;
      if (skip_thread_eligible && tg::thread_eligible(I)) return;
      if (takes(I)) return;
      std::vector<double> ws((size_t)ws_doubles);
      tg::solve_ws_bind(I, ws.data());
      tg::solve_warp(I, 0);  // TG_PHASE runs the 32 lanes of every phase one after the other
    });
  }
  // out[0..count) = indices i < n with flags[i] != 0, ascending; *count = how many
  void select_flagged(const uint8_t* flags, int* out, int* count, int n) {
    int c = 0;
    for (int i = 0; i < n; ++i)
      if (flags[i]) out[c++] = i;
    *count = c;
  }
  void exclusive_scan(const int* in, int* out, int n) {
    int acc = 0;
    for (int i = 0; i < n; ++i) {
      out[i] = acc;
      acc += in[i];
    }
    out[n] = acc;
  }
};

#define TG_BACKEND EmuBackend
#include "../../mrs_uav_trajectory_generation_b200/csrc/tg_capi_impl.hpp"

// measurement hooks exist only in the CUDA build; stubs keep the exported symbol set identical
extern "C" {
int tg_set_profiling(tg_ctx*, int) { return TG_OK; }
int tg_get_profile(tg_ctx*, int, char* names, int names_cap, double*, long long*, long long*) {
  if (names && names_cap > 0) names[0] = 0;
  return 0;
}
double tg_measure_fp64_peak(tg_ctx*, int) { return 0.0; }
}
