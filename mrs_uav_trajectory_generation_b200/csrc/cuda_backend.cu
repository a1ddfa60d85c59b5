// cuda_backend.cu -- the CUDA side of libtg_b200.so: kernels wrapping the functors of tg_kernels.cuh, the CudaBackend
// used by tg_pipeline.hpp, and (through tg_capi_impl.hpp) the C ABI of include/tg_b200.h.
// Build: nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false (see build.py).
// -fmad=false is part of the numeric contract: the reference is built for baseline x86-64 (no FMA contraction,
// CMakeLists.txt:4-12) and 1e-9 coefficient parity needs the same roundings.
#include <cuda_runtime.h>

#include <cub/device/device_scan.cuh>
#include <cub/device/device_select.cuh>
#include <cub/iterator/counting_input_iterator.cuh>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <utility>
#include <vector>
#include <stdexcept>
#include <string>
#include <typeinfo>

#define TG_VERSION_STRING "tg_b200 0.1.0 (sm_100a, fp64, -fmad=false)"

#ifndef TG_JT_REFILL_REP
#define TG_JT_REFILL_REP 2
#endif
#include "tg_kernels.cuh"

namespace {

#define TG_CUDA_CHECK(expr)                                                                              \
  do {                                                                                                   \
    cudaError_t e__ = (expr);                                                                            \
    if (e__ != cudaSuccess)                                                                              \
      throw std::runtime_error(std::string(#expr) + ": " + cudaGetErrorString(e__));                     \
  } while (0)

// one thread per work item.  F::kMinBlocks (optional): resident 128-thread blocks per SM the register allocation must allow
template <class F, class = void>
struct MinBlocksOf {
  static constexpr int value = 1;
};
template <class F>
struct MinBlocksOf<F, decltype((void)F::kMinBlocks)> {
  static constexpr int value = F::kMinBlocks;
};
template <class F>
__global__ void __launch_bounds__(128, MinBlocksOf<F>::value) k_for_each(const F f, const size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) f(i);
}

// one warp per work item (four warps per CTA); f(item, lane) separates its phases with __syncwarp() (TG_PHASE)
template <class F>
__global__ void __launch_bounds__(128) k_for_each_warp(const F f, const size_t n) {
  const size_t w = (size_t)blockIdx.x * 4 + (threadIdx.x >> 5);
  if (w < n) f(w, (int)(threadIdx.x & 31));  // warp-uniform
}

// one thread per work item with F::kScratch doubles of per-thread scratch in dynamic shared memory, element i of thread
// t at smem[i * blockDim.x + t] (bank-conflict free; used by the Jenkins-Traub work arrays)
static_assert(TG_WARR_DEVICE_STRIDE == 32, "strided scratch kernels run one warp per block");
template <class F>
__global__ void __launch_bounds__(TG_WARR_DEVICE_STRIDE) k_for_each_scratch(const F f, const size_t n) {
  extern __shared__ double smem[];
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) f(i, smem + threadIdx.x, (int)blockDim.x);
}

// one warp per solve instance, grid-stride over instances; the per-warp workspace lives in dynamic shared memory
// (or in a global slab when it does not fit).  Phases are separated by __syncwarp() (see tg_solve.cuh).
constexpr int kSolveWarps = 4;
template <class D>
__global__ void __launch_bounds__(kSolveWarps * 32) k_solve(const D desc, const size_t inst_begin, const size_t n_inst, const int ws_doubles, double* __restrict__ gws,
                                                            const int skip_oct_ws, const bool skip_thread) {
  extern __shared__ double smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const size_t gw = (size_t)blockIdx.x * kSolveWarps + warp, nw = (size_t)gridDim.x * kSolveWarps;
  double* ws = gws ? gws + gw * (size_t)ws_doubles : smem + (size_t)warp * ws_doubles;
  for (size_t inst = inst_begin + gw; inst < n_inst; inst += nw) {
    tg::SolveInst I;
    if (!desc.instance(inst, I)) continue;  // warp-uniform
    if (skip_thread && tg::thread_eligible(I)) continue;  // k_solve_thread has taken it
    // skip_oct_ws > 0: the octet kernel has taken every instance it can take (same test as in k_solve_oct)
    if (skip_oct_ws > 0 && tg::octet_eligible(I) && tg::octet_ws_doubles(I.S, I.np) <= skip_oct_ws) continue;
    tg::solve_ws_bind(I, ws);
    tg::solve_warp(I, lane);
  }
}

// Four solve instances per warp, eight lanes each (tg_solve_octet.cuh); instances the octet routine cannot take are left
// to k_solve.  One warp per CTA.  uslab: per-CTA slab of 4 * u_cap rows of kOctURow doubles (the U rows between the
// elimination and the back substitution; L2 resident).
template <class D>
__global__ void __launch_bounds__(32) k_solve_oct(const D desc, const size_t inst_begin, const size_t n_inst, const int oct_ws_doubles, const int u_cap,
                                                  double* __restrict__ uslab, const bool skip_thread) {
  extern __shared__ double smem[];
  const int lane = threadIdx.x & 31, oct = lane >> 3;
  const size_t gw = blockIdx.x, nw = gridDim.x;
  double* urows = uslab + ((size_t)blockIdx.x * 4 + oct) * (size_t)u_cap * tg::kOctURow;
  for (size_t base = inst_begin + gw * 4; base < n_inst; base += nw * 4) {
    tg::SolveInst I;
    const size_t inst = base + oct;
    bool ok = inst < n_inst && desc.instance(inst, I);
    if (ok && skip_thread && tg::thread_eligible(I)) ok = false;  // k_solve_thread has taken it
    if (ok && !(tg::octet_eligible(I) && tg::octet_ws_doubles(I.S, I.np) <= oct_ws_doubles && I.np <= u_cap)) ok = false;
    if (!ok) {
      I.S = 0; I.np = 0; I.hbw = tg::kOctHbw; I.dp_out = nullptr; I.coef_out = nullptr; I.cost_out = nullptr;
    }
    const int nmax = __reduce_max_sync(0xffffffffu, I.np);
    if (nmax > 0) {
      tg::octet_ws_bind(I, smem + (size_t)oct * oct_ws_doubles, urows);
      tg::solve_octets(&I, lane, nmax);
    }
    __syncwarp();
  }
}

// Exact per-segment maxima of quantity Q for the entries of a device work list, by PERSISTENT warps with lane refill
// (profiles/r02_extrema_refill.md).  The plain launch (k_for_each_scratch<ExtremaRawFn<Q>>) gives every thread one
// polynomial: Jenkins-Traub run times vary 3x between polynomials, so a warp idles 60 % of its lanes while the slowest
// finishes, and the launch lasts as long as its slowest warp.  Here a lane whose polynomial is done stores its maximum
// and takes the next work item at once; every pass of the loop executes one block of the stage machine (tg_poly.cuh),
// the one most lanes are waiting for.  Lanes are independent, so the order of execution cannot change a result.
// Loading a work item is kept out of line (the unrolled convolution needs ~100 registers that the stage machine's loop should
// not pay for): candidates t = 0 and t = T, the polynomial into the lane's shared-memory array p (decreasing powers, trimmed
// as findRootsJenkinsTraub does, rpoly_ak1.cpp:76-120, 174-180).  Returns the degree to iterate on (-1: nothing) and the
// best magnitude so far.
struct RefillLoad {
  int degree;
  double best;
};
template <int Q>
__device__ __noinline__ RefillLoad extrema_load_item(const double* __restrict__ c, const double T, double* __restrict__ p_lane) {
  typedef tg::QuantityJob<Q> Job;
  constexpr int M = Job::M;
  typename Job::Sink sink{c, T, TG_DBL_LOWEST};
  RefillLoad r{-1, TG_DBL_LOWEST};
  if (0.0 > T) return r;
  sink.consider(0.0);
  sink.consider(T);
  // same sums as QuantityJob<Q>::poly, with rolled loops over arrays in local memory (this path runs once per polynomial;
  // unrolled it would set the register count of the whole kernel)
  double ci[M + 1];
  if constexpr (Job::kPair) {
    constexpr int n_d = TG_N - Job::kDeriv, n_dd = n_d - 1, len = n_d + n_dd - 1;
#pragma unroll 1
    for (int i = 0; i < len; ++i) ci[i] = 0.0;
#pragma unroll 1
    for (int dim = 0; dim < 2; ++dim) {
      const double* cc = c + dim * TG_N;
      double dc[n_d], ddc[n_dd];
#pragma unroll 1
      for (int jx = 0; jx < n_d; ++jx) dc[jx] = cc[jx + Job::kDeriv] * tg::bcoef(Job::kDeriv, jx + Job::kDeriv);
#pragma unroll 1
      for (int jx = 0; jx < n_dd; ++jx) ddc[jx] = cc[jx + Job::kDeriv + 1] * tg::bcoef(Job::kDeriv + 1, jx + Job::kDeriv + 1);
#pragma unroll 1
      for (int i = 0; i < len; ++i) {
        double cv = 0.0;
        const int data_idx = i - n_dd + 1;
        const int lower = (0 > -data_idx) ? 0 : -data_idx, upper = (n_dd < n_d - data_idx) ? n_dd : n_d - data_idx;
#pragma unroll 1
        for (int kidx = lower; kidx < upper; ++kidx) cv = cv + ddc[n_dd - 1 - kidx] * dc[data_idx + kidx];
        ci[i] = ci[i] + cv;
      }
    }
  } else {
    const double* cc = c + Job::kD0 * TG_N;
#pragma unroll 1
    for (int jx = 0; jx <= M; ++jx) ci[jx] = cc[jx + Job::kDeriv + 1] * tg::bcoef(Job::kDeriv + 1, jx + Job::kDeriv + 1);
  }
  int last = -1;
#pragma unroll 1
  for (int i = 0; i <= M; i++)
    if (tg::dabs(ci[i]) >= TG_DBL_MIN) last = i;
  if (last >= 1) {
    int low = last;
#pragma unroll 1
    for (int i = M; i >= 0; i--)
      if (i <= last && ci[i] != 0.0) low = i;
    for (int z = 0; z < low; ++z) sink(0.0, 0.0);
    r.degree = last - low;
#pragma unroll 1
    for (int i = 0; i <= M; i++) {
      const int dst = last - i;
      if (dst >= 0 && dst <= r.degree) p_lane[dst * 32] = ci[i];
    }
  }
  r.best = sink.best;
  return r;
}

template <int Q>
__global__ void __launch_bounds__(32) k_extrema_refill(const double* __restrict__ coef, const double* __restrict__ times, double* __restrict__ maxima,
                                                       const int* __restrict__ work, const int* __restrict__ n_dev, const int n_max, int* __restrict__ counter,
                                                       unsigned long long* __restrict__ flop_counter) {
  extern __shared__ double smem[];
  typedef tg::QuantityJob<Q> Job;
  typedef typename Job::Sink Sink;
  constexpr int M = Job::M;
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const int n = n_dev ? min(n_max, *n_dev) : n_max;
  double* scratch = smem + lane;
  double svk[M + 1], tmp[M + 1];
  tg::JtMachine m;
  m.p = tg::WArr{scratch, 32};
  m.qp = tg::WArr{scratch + (size_t)(M + 1) * 32, 32};
  m.K = tg::WArr{scratch + (size_t)2 * (M + 1) * 32, 32};
  m.qk = tg::WArr{scratch + (size_t)3 * (M + 1) * 32, 32};
  m.svk = svk;
  m.tmp = tmp;
  m.state = tg::JtMachine::kDone;
  Sink sink{nullptr, 0.0, TG_DBL_LOWEST};
  size_t seg = 0;
  bool has = false, exhausted = false;
  m.fl = 0;
  // counted floating-point operations of this lane (stage-machine blocks + the item loads), reported when profiling
  constexpr int kLoadFlops = Job::kPair ? 4 * (TG_N - Job::kDeriv) * (TG_N - Job::kDeriv - 1) + 8 * (TG_N - Job::kDeriv) : (M + 1) + 4 * (TG_N - Job::kDeriv);
  unsigned long long lane_fl = 0;
  for (;;) {
    const bool idle = (m.state == tg::JtMachine::kDone);
    const unsigned idle_mask = __ballot_sync(full, idle);
    if (idle_mask) {
      if (idle && has) {
        maxima[seg * 9 + Q] = sink.best;
        has = false;
        lane_fl += (unsigned long long)(m.fl + kLoadFlops);
        m.fl = 0;
      }
      if (!exhausted) {
        const int cnt = __popc(idle_mask), leader = __ffs(idle_mask) - 1;
        int base = 0;
        if (lane == leader) base = atomicAdd(counter, cnt);
        base = __shfl_sync(full, base, leader);
        if (base + cnt >= n) exhausted = true;
        if (idle) {
          const int my = base + __popc(idle_mask & ((1u << lane) - 1u));
          if (my < n) {
            seg = work ? (size_t)work[my] : (size_t)my;
            const double* c = coef + seg * TG_D * TG_N;
            const double T = times[seg];
            const RefillLoad ld = extrema_load_item<Q>(c, T, scratch);
            sink = Sink{c, T, ld.best};
            has = true;
            if (ld.degree >= 0) m.begin(ld.degree);
          }
        }
      }
    }
    const int st = m.state;
    const unsigned same = __match_any_sync(full, st);
    const unsigned key = (st == tg::JtMachine::kDone) ? 0u : (((unsigned)__popc(same) << 8) | (unsigned)(st + 1));
    const unsigned win = __reduce_max_sync(full, key);
    if (win == 0u) {
      if (exhausted && !__any_sync(full, has)) break;
      continue;
    }
    const int cur = (int)(win & 0xffu) - 1;
    if (st == cur) {
#pragma unroll 1
      for (int rep = 0; rep < TG_JT_REFILL_REP && m.state == cur; ++rep) m.step(cur, sink, nullptr);
    }
  }
  if (flop_counter) {
    for (int o = 16; o > 0; o >>= 1) lane_fl += __shfl_xor_sync(full, lane_fl, o);
    if (lane == 0) atomicAdd(flop_counter, lane_fl);
  }
}

// One thread per solve instance (tg_solve_thread.cuh), grid-stride over instances.  slab: rows_cap rows of kThrRow doubles
// per thread, interleaved by thread (element e of row r of thread t at slab[(r * kThrRow + e) * nthreads + t]) so that the
// accesses of a warp coalesce.
constexpr int kThreadSolveCta = 128;
template <class D>
__global__ void __launch_bounds__(kThreadSolveCta) k_solve_thread(const D desc, const size_t inst_begin, const size_t inst_end, double* __restrict__ slab) {
  const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nthreads = (size_t)gridDim.x * blockDim.x;
  for (size_t inst = inst_begin + tid; inst < inst_end; inst += nthreads) {
    tg::SolveInst I;
    if (!desc.instance(inst, I)) continue;
    if (!tg::thread_eligible(I)) continue;
    tg::solve_thread(I, slab + tid, nthreads);
  }
}

// FP64 pipe peak probes (the roofline denominator for this path; MEASURED_PEAKS.json has HBM and bf16 only).
// mode 0: DFMA chains; mode 1: DMUL+DADD pairs as generated under -fmad=false (what the product kernels issue).
__global__ void __launch_bounds__(256) k_fp64_peak(double* out, int iters, int mode) {
  double a0 = 1.0 + threadIdx.x * 1e-9, a1 = a0 + 1e-3, a2 = a0 + 2e-3, a3 = a0 + 3e-3, a4 = a0 + 4e-3, a5 = a0 + 5e-3, a6 = a0 + 6e-3,
         a7 = a0 + 7e-3;
  const double m = 1.0 - 1e-12, c = 1e-13;
  if (mode == 0) {
    for (int i = 0; i < iters; ++i) {
      a0 = __fma_rn(a0, m, c); a1 = __fma_rn(a1, m, c); a2 = __fma_rn(a2, m, c); a3 = __fma_rn(a3, m, c);
      a4 = __fma_rn(a4, m, c); a5 = __fma_rn(a5, m, c); a6 = __fma_rn(a6, m, c); a7 = __fma_rn(a7, m, c);
    }
  } else {
    for (int i = 0; i < iters; ++i) {
      a0 = __dadd_rn(__dmul_rn(a0, m), c); a1 = __dadd_rn(__dmul_rn(a1, m), c); a2 = __dadd_rn(__dmul_rn(a2, m), c);
      a3 = __dadd_rn(__dmul_rn(a3, m), c); a4 = __dadd_rn(__dmul_rn(a4, m), c); a5 = __dadd_rn(__dmul_rn(a5, m), c);
      a6 = __dadd_rn(__dmul_rn(a6, m), c); a7 = __dadd_rn(__dmul_rn(a7, m), c);
    }
  }
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
}

}  // namespace

struct CudaBackend {
  int device = 0;
  int sm_count = 148;
  size_t smem_optin = 0, smem_per_sm = 0;
  bool force_general_solve = false;
  int oct_reg_warps = 8;  // resident single-warp CTAs of k_solve_oct per SM as far as registers allow
  cudaStream_t stream = nullptr;
  cudaStream_t copy_stream = nullptr;  // d2h_overlapped
  cudaEvent_t copy_ready = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  static constexpr int kSideStreams = 9;
  cudaStream_t side[kSideStreams] = {};
  cudaEvent_t ev_fork = nullptr, ev_side[kSideStreams] = {};
  void* scan_tmp = nullptr;
  size_t scan_tmp_bytes = 0;
  static constexpr size_t kStageBytes = 64 * 1024;
  static constexpr size_t kBigStageBytes = 16u << 20;
  void* h_stage_big = nullptr;
  bool use_big_stage = std::getenv("TG_BIG_STAGE") ? std::atoi(std::getenv("TG_BIG_STAGE")) != 0 : true;
  void* h_stage = nullptr;                       // pinned landing block of the small device-to-host reads
  double* solve_slab = nullptr;  // U rows of the octet kernel
  size_t solve_slab_doubles = 0;
  double* gen_slab = nullptr;    // workspaces of the warp-per-instance kernel for very long paths
  size_t gen_slab_doubles = 0;
  // optional per-kernel timing (bench.py roofline leg): events around every launch, accumulated per functor type
  bool profiling = false;
  cudaEvent_t pev0 = nullptr, pev1 = nullptr;
  struct ProfEntry { double ms = 0; long long launches = 0; long long items = 0; };
  std::map<std::string, ProfEntry> prof;

  explicit CudaBackend(int dev) : device(dev) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) throw std::runtime_error("no CUDA device available (libtg_b200 has no CPU fallback)");
    if (dev < 0 || dev >= n) throw std::runtime_error("CUDA device index out of range");
    TG_CUDA_CHECK(cudaSetDevice(dev));
    cudaDeviceProp prop;
    TG_CUDA_CHECK(cudaGetDeviceProperties(&prop, dev));
    sm_count = prop.multiProcessorCount;
    smem_optin = prop.sharedMemPerBlockOptin;
    smem_per_sm = prop.sharedMemPerMultiprocessor;
    force_general_solve = std::getenv("TG_NO_OCTET") != nullptr;
    if (std::getenv("TG_L2_PERSIST") != nullptr && prop.persistingL2CacheMaxSize > 0) {  // opt-in: measured slower (profiles/r02_solve_thread.md)
      size_t want = (size_t)prop.persistingL2CacheMaxSize;
      if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want) == cudaSuccess) {
        cudaDeviceGetLimit(&l2_persist_bytes, cudaLimitPersistingL2CacheSize);
        l2_window_max = (size_t)prop.accessPolicyMaxWindowSize;
      } else {
        cudaGetLastError();
      }
    }
    {
      int a = 0, b = 0;
      TG_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&a, k_solve_oct<tg::SolveProblemDesc>, 32, 0));
      TG_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, k_solve_oct<tg::SolveSweepDesc>, 32, 0));
      oct_reg_warps = std::max(1, std::min(a, b));
      if (const char* e = std::getenv("TG_OCT_WARPS")) oct_reg_warps = std::max(1, std::min(oct_reg_warps, std::atoi(e)));  // experiments
    }
    TG_CUDA_CHECK(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    TG_CUDA_CHECK(cudaHostAlloc(&h_stage, kStageBytes, cudaHostAllocDefault));
    TG_CUDA_CHECK(cudaEventCreate(&ev0));
    TG_CUDA_CHECK(cudaEventCreate(&ev1));
    TG_CUDA_CHECK(cudaEventCreate(&pev0));
    TG_CUDA_CHECK(cudaEventCreate(&pev1));
    TG_CUDA_CHECK(cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming));
    for (int i = 0; i < kSideStreams; ++i) {
      TG_CUDA_CHECK(cudaStreamCreateWithFlags(&side[i], cudaStreamNonBlocking));
      TG_CUDA_CHECK(cudaEventCreateWithFlags(&ev_side[i], cudaEventDisableTiming));
    }
  }
  ~CudaBackend() {
    cudaSetDevice(device);
    if (scan_tmp) cudaFree(scan_tmp);
    if (h_stage) cudaFreeHost(h_stage);
    if (h_stage_big) cudaFreeHost(h_stage_big);
    for (PinnedChunk& c : pinned_chunks) {
      if (pinned_state) cudaFreeHost(c.p);
      else delete[] c.p;
    }
    if (solve_slab) cudaFree(solve_slab);
    if (thread_slab) cudaFree(thread_slab);
    if (gen_slab) cudaFree(gen_slab);
    if (ev0) cudaEventDestroy(ev0);
    if (ev1) cudaEventDestroy(ev1);
    if (pev0) cudaEventDestroy(pev0);
    if (pev1) cudaEventDestroy(pev1);
    for (int i = 0; i < kSideStreams; ++i) {
      if (side[i]) cudaStreamDestroy(side[i]);
      if (ev_side[i]) cudaEventDestroy(ev_side[i]);
    }
    if (ev_fork) cudaEventDestroy(ev_fork);
    if (copy_ready) cudaEventDestroy(copy_ready);
    if (copy_stream) cudaStreamDestroy(copy_stream);
    if (stream) cudaStreamDestroy(stream);
  }
  CudaBackend(const CudaBackend&) = delete;
  CudaBackend& operator=(const CudaBackend&) = delete;

  void bind() { TG_CUDA_CHECK(cudaSetDevice(device)); }
  void* dev_alloc(size_t n) {
    bind();
    void* p = nullptr;
    TG_CUDA_CHECK(cudaMalloc(&p, n));
    return p;
  }
  void dev_free(void* p) {
    cudaSetDevice(device);
    cudaFree(p);
  }
  void h2d(void* d, const void* s, size_t n) {
    if (n) TG_CUDA_CHECK(cudaMemcpyAsync(d, s, n, cudaMemcpyHostToDevice, stream));
  }
  // small read-backs (counters, list sizes) land in a pinned staging block: a pageable destination makes the driver stage and block
  void d2h(void* d, const void* s, size_t n) {
    if (n && n <= kStageBytes && h_stage) {
      TG_CUDA_CHECK(cudaMemcpyAsync(h_stage, s, n, cudaMemcpyDeviceToHost, stream));
      wait_stream();
      std::memcpy(d, h_stage, n);
      return;
    }
    // medium read-backs (per-problem state of a group: 72 B x 65 536 paths, offsets, unknown counts) go through a larger pinned block as
    // well: a pageable destination measured ~2 ms per 4.7 MB here with the device idle, the staged copy + memcpy a quarter of that.
    // Large copies (the caller's output buffers, which a caller who cares page-locks) go straight through.
    if (n && n <= kBigStageBytes && use_big_stage) {
      if (!h_stage_big) TG_CUDA_CHECK(cudaHostAlloc(&h_stage_big, kBigStageBytes, cudaHostAllocDefault));
      TG_CUDA_CHECK(cudaMemcpyAsync(h_stage_big, s, n, cudaMemcpyDeviceToHost, stream));
      wait_stream();
      std::memcpy(d, h_stage_big, n);
      return;
    }
    if (n) TG_CUDA_CHECK(cudaMemcpyAsync(d, s, n, cudaMemcpyDeviceToHost, stream));
    wait_stream();
  }
  // dynamic shared memory opt-in, once per kernel and size (not on every launch)
  // (the attribute belongs to the function on a device, not to a context: the record is shared by all contexts of the process and
  // only ever raised, otherwise a second context would lower what the first one relies on)
  void allow_smem(const void* fn, size_t smem) {
    static std::mutex mu;
    static std::map<std::pair<int, const void*>, size_t> allowed;
    std::lock_guard<std::mutex> lock(mu);
    size_t& have = allowed[std::make_pair(device, fn)];
    if (smem <= have) return;
    TG_CUDA_CHECK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    have = smem;
  }
  void d2d(void* d, const void* s, size_t n) {
    if (n) TG_CUDA_CHECK(cudaMemcpyAsync(d, s, n, cudaMemcpyDeviceToDevice, stream));
  }
  void dev_memset(void* d, int v, size_t n) {
    if (n) TG_CUDA_CHECK(cudaMemsetAsync(d, v, n, stream));
  }
  void sync() { wait_stream(); }
  // Page-locked host memory for read-backs the host loops over right away (the per-problem state of a group): a bump arena kept across
  // calls, reset with the device arenas.  Copying into it needs no staging block and no second memcpy (1 ms per 4.7 MB otherwise).
  struct PinnedChunk { char* p; size_t cap; };
  std::vector<PinnedChunk> pinned_chunks;
  size_t pinned_chunk = 0, pinned_used = 0;
  const bool pinned_state = std::getenv("TG_PINNED_STATE") ? std::atoi(std::getenv("TG_PINNED_STATE")) != 0 : true;
  void pinned_reset() { pinned_chunk = 0; pinned_used = 0; }
  void* pinned_alloc(size_t bytes) {
    bytes = (bytes + 255) & ~(size_t)255;
    for (;;) {
      if (pinned_chunk < pinned_chunks.size()) {
        PinnedChunk& c = pinned_chunks[pinned_chunk];
        if (pinned_used + bytes <= c.cap) {
          void* out = c.p + pinned_used;
          pinned_used += bytes;
          return out;
        }
        ++pinned_chunk;
        pinned_used = 0;
        continue;
      }
      PinnedChunk c;
      c.cap = std::max(bytes, (size_t)8 << 20);
      if (pinned_state) TG_CUDA_CHECK(cudaHostAlloc((void**)&c.p, c.cap, cudaHostAllocDefault));
      else c.p = new char[c.cap];  // TG_PINNED_STATE=0 (A/B measurements): pageable memory, staged copies
      pinned_chunks.push_back(c);
    }
  }
  void d2h_pinned(void* d, const void* s, size_t n) {
    if (!pinned_state) { d2h(d, s, n); return; }
    if (n) TG_CUDA_CHECK(cudaMemcpyAsync(d, s, n, cudaMemcpyDeviceToHost, stream));
    wait_stream();
  }
  // every host wait on the compute stream goes through here; TG_TRACE_HOST=1 accumulates the time spent waiting, so that
  // (call time - wait time) = host work during which the device may sit idle (printed per batch call by tg_optimize_batch)
  double wait_s = 0.0;
  long long waits = 0;
  const bool trace_host = std::getenv("TG_TRACE_HOST") != nullptr;
  void wait_stream() {
    if (!trace_host) {
      TG_CUDA_CHECK(cudaStreamSynchronize(stream));
      return;
    }
    const auto t0 = std::chrono::steady_clock::now();
    TG_CUDA_CHECK(cudaStreamSynchronize(stream));
    wait_s += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    ++waits;
  }
  // Device-to-host copy that does not hold up the compute stream: it waits for what has been enqueued so far, then runs on its own
  // stream while later launches proceed (a real overlap needs a pinned destination; a pageable one still gives the right bytes).
  // copy_join() waits for every such copy.
  void d2h_overlapped(void* d, const void* s, size_t n) {
    if (!n) return;
    if (!copy_stream) {
      TG_CUDA_CHECK(cudaStreamCreateWithFlags(&copy_stream, cudaStreamNonBlocking));
      TG_CUDA_CHECK(cudaEventCreateWithFlags(&copy_ready, cudaEventDisableTiming));
    }
    TG_CUDA_CHECK(cudaEventRecord(copy_ready, stream));
    TG_CUDA_CHECK(cudaStreamWaitEvent(copy_stream, copy_ready, 0));
    TG_CUDA_CHECK(cudaMemcpyAsync(d, s, n, cudaMemcpyDeviceToHost, copy_stream));
  }
  void copy_join() {
    if (copy_stream) TG_CUDA_CHECK(cudaStreamSynchronize(copy_stream));
  }
  void timer_start() {
    bind();
    TG_CUDA_CHECK(cudaEventRecord(ev0, stream));
  }
  double timer_stop() {
    TG_CUDA_CHECK(cudaEventRecord(ev1, stream));
    TG_CUDA_CHECK(cudaEventSynchronize(ev1));
    float ms = 0.f;
    TG_CUDA_CHECK(cudaEventElapsedTime(&ms, ev0, ev1));
    return (double)ms;
  }

  template <class F>
  void for_each(size_t n, const F& f) {
    if (n == 0) return;
    const unsigned block = 128;
    const size_t grid = (n + block - 1) / block;
    prof_begin();
    k_for_each<F><<<(unsigned)grid, block, 0, stream>>>(f, n);
    TG_CUDA_CHECK(cudaGetLastError());
    prof_end(typeid(F).name(), n);
  }
  template <class F>
  void for_each_warp(size_t n, const F& f) {
    if (n == 0) return;
    const size_t grid = (n + 3) / 4;
    prof_begin();
    k_for_each_warp<F><<<(unsigned)grid, 128, 0, stream>>>(f, n);
    TG_CUDA_CHECK(cudaGetLastError());
    prof_end(typeid(F).name(), n);
  }
  template <class F>
  void for_each_scratch(size_t n, const F& f) {
    if (n == 0) return;
    const unsigned block = TG_WARR_DEVICE_STRIDE;
    const size_t grid = (n + block - 1) / block;
    const size_t smem = (size_t)F::kScratch * sizeof(double) * block;
    allow_smem((const void*)k_for_each_scratch<F>, smem);
    prof_begin();
    k_for_each_scratch<F><<<(unsigned)grid, block, smem, stream>>>(f, n);
    TG_CUDA_CHECK(cudaGetLastError());
    prof_end(typeid(F).name(), n);
  }
  // fork / join of n side streams around independent launches (for_each_scratch_on); with per-kernel profiling on, everything
  // stays on the main stream so that the event pairs measure single kernels
  void fork(int n) {
    if (profiling) return;
    TG_CUDA_CHECK(cudaEventRecord(ev_fork, stream));
    for (int i = 0; i < n && i < kSideStreams; ++i) TG_CUDA_CHECK(cudaStreamWaitEvent(side[i], ev_fork, 0));
  }
  void join(int n) {
    if (profiling) return;
    for (int i = 0; i < n && i < kSideStreams; ++i) {
      TG_CUDA_CHECK(cudaEventRecord(ev_side[i], side[i]));
      TG_CUDA_CHECK(cudaStreamWaitEvent(stream, ev_side[i], 0));
    }
  }
  template <class F>
  void for_each_scratch_on(int k, size_t n, const F& f) {
    if (profiling) return for_each_scratch(n, f);
    if (n == 0) return;
    const unsigned block = TG_WARR_DEVICE_STRIDE;
    const size_t grid = (n + block - 1) / block;
    const size_t smem = (size_t)F::kScratch * sizeof(double) * block;
    allow_smem((const void*)k_for_each_scratch<F>, smem);
    k_for_each_scratch<F><<<(unsigned)grid, block, smem, side[k % kSideStreams]>>>(f, n);
    TG_CUDA_CHECK(cudaGetLastError());
  }
  // exact maxima of quantity Q for a device work list through the persistent refill kernel, on side stream k
  int refill_ctas_per_sm[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  unsigned long long* d_jt_flops = nullptr;
  bool use_refill = std::getenv("TG_JT_REFILL") != nullptr;  // opt-in: measured no faster than one polynomial per thread (profiles/r02_extrema_refill.md)
  template <int Q>
  void extrema_refill(int k, size_t n_max, const double* coef, const double* times, double* maxima, const int* work, const int* n_dev, int* counter) {
    if (n_max == 0) return;
    if (!use_refill) {
      tg::ExtremaRawFn<Q> f{coef, times, maxima, work, n_dev};
      if (!profiling) return for_each_scratch_on(k, n_max, f);
      // profile step: the same launch with its executed operations counted (`items` of the "jt:" entries = FP64 operations)
      if (!d_jt_flops) TG_CUDA_CHECK(cudaMalloc(&d_jt_flops, sizeof(unsigned long long)));
      TG_CUDA_CHECK(cudaMemsetAsync(d_jt_flops, 0, sizeof(unsigned long long), stream));
      const std::string name = std::string("jt:ExtremaRawFn<") + std::to_string(Q) + ">";
      const size_t smem = (size_t)tg::ExtremaRawFn<Q>::kScratch * sizeof(double) * TG_WARR_DEVICE_STRIDE;
      allow_smem((const void*)k_for_each_scratch<tg::ExtremaRawCountFn<Q>>, smem);
      prof_begin();
      k_for_each_scratch<tg::ExtremaRawCountFn<Q>><<<(unsigned)((n_max + 31) / 32), TG_WARR_DEVICE_STRIDE, smem, stream>>>(tg::ExtremaRawCountFn<Q>{f, d_jt_flops}, n_max);
      TG_CUDA_CHECK(cudaGetLastError());
      prof_end(name.c_str(), 0);
      unsigned long long fl = 0;
      TG_CUDA_CHECK(cudaMemcpy(&fl, d_jt_flops, sizeof(fl), cudaMemcpyDeviceToHost));
      prof[name].items += (long long)fl;
      return;
    }
    constexpr int M = tg::QuantityJob<Q>::M;
    const size_t smem = (size_t)4 * (M + 1) * sizeof(double) * 32;
    if (refill_ctas_per_sm[Q] == 0) {
      allow_smem((const void*)k_extrema_refill<Q>, smem);
      int a = 0;
      TG_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&a, k_extrema_refill<Q>, 32, smem));
      refill_ctas_per_sm[Q] = std::max(a, 1);
    }
    const size_t grid = std::min((n_max + 31) / 32, (size_t)sm_count * refill_ctas_per_sm[Q]);
    cudaStream_t st = profiling ? stream : side[k % kSideStreams];
    if (profiling) {
      if (!d_jt_flops) TG_CUDA_CHECK(cudaMalloc(&d_jt_flops, sizeof(unsigned long long)));
      TG_CUDA_CHECK(cudaMemsetAsync(d_jt_flops, 0, sizeof(unsigned long long), stream));
    }
    prof_begin();
    k_extrema_refill<Q><<<(unsigned)grid, 32, smem, st>>>(coef, times, maxima, work, n_dev, (int)n_max, counter, profiling ? d_jt_flops : nullptr);
    TG_CUDA_CHECK(cudaGetLastError());
    if (profiling) {
      // `items` of these entries = counted floating-point operations (the kernel adds them up per polynomial)
      const std::string name = std::string("refill:ExtremaRawFn<") + std::to_string(Q) + ">";
      prof_end(name.c_str(), 0);
      unsigned long long fl = 0;
      TG_CUDA_CHECK(cudaMemcpy(&fl, d_jt_flops, sizeof(fl), cudaMemcpyDeviceToHost));
      prof[name].items += (long long)fl;
    }
  }
  void prof_begin() {
    if (profiling) TG_CUDA_CHECK(cudaEventRecord(pev0, stream));
  }
  void prof_end(const char* name, size_t items) {
    if (!profiling) return;
    TG_CUDA_CHECK(cudaEventRecord(pev1, stream));
    TG_CUDA_CHECK(cudaEventSynchronize(pev1));
    float ms = 0.f;
    TG_CUDA_CHECK(cudaEventElapsedTime(&ms, pev0, pev1));
    ProfEntry& e = prof[name];
    e.ms += ms;
    e.launches += 1;
    e.items += (long long)items;
  }
  // TFLOP/s of the FP64 pipe: mode 0 DFMA (2 flop), mode 1 DMUL+DADD (2 flop in two instructions)
  double fp64_peak_tflops(int mode) {
    bind();
    const int blocks = sm_count * 16, threads = 256, iters = 1 << 14;
    double* out = nullptr;
    TG_CUDA_CHECK(cudaMalloc(&out, sizeof(double) * (size_t)blocks * threads));
    double best = 0.0;
    for (int rep = 0; rep < 4; ++rep) {
      TG_CUDA_CHECK(cudaEventRecord(pev0, stream));
      k_fp64_peak<<<blocks, threads, 0, stream>>>(out, iters, mode);
      TG_CUDA_CHECK(cudaEventRecord(pev1, stream));
      TG_CUDA_CHECK(cudaEventSynchronize(pev1));
      float ms = 0.f;
      TG_CUDA_CHECK(cudaEventElapsedTime(&ms, pev0, pev1));
      const double flop = 2.0 * 8.0 * (double)iters * (double)blocks * threads;
      if (rep > 0) best = std::max(best, flop / (ms * 1e-3) / 1e12);
    }
    cudaFree(out);
    return best;
  }

  // resident warps per SM of the octet kernel for a given per-octet workspace (0: not eligible)
  int octet_warps_per_sm(int oct_ws_doubles) const {
    if (oct_ws_doubles <= 0 || force_general_solve) return 0;
    const size_t smem = (size_t)4 * oct_ws_doubles * sizeof(double);
    if (smem > smem_optin) return 0;
    const int by_smem = (int)(smem_per_sm / (smem + 1024));  // 1 KB per CTA is reserved by the system
    return std::min(by_smem, oct_reg_warps);
  }
  // Class of a problem's solve workspace: the pipeline cuts a batch into runs of equal class (tg_pipeline.hpp
  // make_buckets) so that each run is launched with its own shared-memory size and every instance of a run goes to the
  // same kernel.  >0: octet kernel, that many warps per SM; 0: warp-per-instance kernel in shared memory; -1: warp-per-
  // instance kernel with its workspace in the global slab.
  int solve_class(int ws_doubles, int oct_ws_doubles) const {
    const int w = octet_warps_per_sm(oct_ws_doubles);
    if (w >= 1) return w;
    return ((size_t)ws_doubles * sizeof(double) * kSolveWarps <= smem_optin) ? 0 : -1;
  }

  double* slab(size_t doubles) {
    if (doubles > solve_slab_doubles) {
      if (solve_slab) {
        TG_CUDA_CHECK(cudaStreamSynchronize(stream));
        TG_CUDA_CHECK(cudaFree(solve_slab));
      }
      TG_CUDA_CHECK(cudaMalloc(&solve_slab, doubles * sizeof(double)));
      solve_slab_doubles = doubles;
    }
    return solve_slab;
  }

  bool skip_thread_eligible = false;  // set by the pipeline after solve_thread(): the older kernels leave those instances alone
  size_t l2_persist_bytes = 0, l2_window_max = 0;  // persisting L2 set-aside and the largest access-policy window (0: unsupported / disabled)
  double* thread_slab = nullptr;
  size_t thread_slab_doubles = 0;
  int thread_ctas_per_sm = 0;
  // Thread-per-instance solve of every eligible instance of [inst_begin, inst_end); rows_cap = slab rows per instance.
  template <class D>
  void solve_thread(size_t inst_begin, size_t inst_end, int rows_cap, const D& desc) {
    if (inst_end <= inst_begin) return;
    const size_t n_inst = inst_end - inst_begin;
    if (thread_ctas_per_sm == 0) {
      int a = 0;
      TG_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&a, k_solve_thread<tg::SolveProblemDesc>, kThreadSolveCta, 0));
      thread_ctas_per_sm = std::max(a, 1);
      if (const char* e = std::getenv("TG_THREAD_CTAS")) thread_ctas_per_sm = std::max(1, std::atoi(e));
    }
    const size_t blocks_needed = (n_inst + kThreadSolveCta - 1) / kThreadSolveCta;
    const size_t grid = std::min(blocks_needed, (size_t)sm_count * thread_ctas_per_sm);
    const size_t need = grid * kThreadSolveCta * (size_t)rows_cap * tg::kThrRow;
    if (need > thread_slab_doubles) {
      if (thread_slab) {
        TG_CUDA_CHECK(cudaStreamSynchronize(stream));
        TG_CUDA_CHECK(cudaFree(thread_slab));
      }
      TG_CUDA_CHECK(cudaMalloc(&thread_slab, need * sizeof(double)));
      thread_slab_doubles = need;
    }
    // the slab is written and read back by the same resident threads over and over: keep as much of it as the device allows
    // in the persisting part of L2, so that it does not stream through HBM (profiles/r02_solve_thread.md)
    const bool persist = l2_persist_bytes > 0 && n_inst > (size_t)4096;
    if (persist) {
      cudaStreamAttrValue av;
      std::memset(&av, 0, sizeof(av));
      const size_t bytes = std::min(need * sizeof(double), (size_t)l2_window_max);
      av.accessPolicyWindow.base_ptr = thread_slab;
      av.accessPolicyWindow.num_bytes = bytes;
      av.accessPolicyWindow.hitRatio = (float)std::min(1.0, (double)l2_persist_bytes / (double)bytes);
      av.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
      av.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
      TG_CUDA_CHECK(cudaStreamSetAttribute(stream, cudaStreamAttributeAccessPolicyWindow, &av));
    }
    prof_begin();
    k_solve_thread<D><<<(unsigned)grid, kThreadSolveCta, 0, stream>>>(desc, inst_begin, inst_end, thread_slab);
    TG_CUDA_CHECK(cudaGetLastError());
    prof_end((std::string("thread:") + typeid(D).name()).c_str(), n_inst);
    if (persist) {
      cudaStreamAttrValue av;
      std::memset(&av, 0, sizeof(av));
      av.accessPolicyWindow.num_bytes = 0;
      TG_CUDA_CHECK(cudaStreamSetAttribute(stream, cudaStreamAttributeAccessPolicyWindow, &av));
    }
  }

  // Solves instances [inst_begin, inst_end).  oct_ws_doubles > 0: instances the octet routine can take (half bandwidth
  // 7, workspace <= oct_ws_doubles, at most np_cap unknowns) go to k_solve_oct; `mixed` says that others may be present,
  // which then go to the warp-per-instance kernel (ws_doubles of workspace each).
  template <class D>
  void solve(size_t inst_begin, size_t inst_end, int ws_doubles, int oct_ws_doubles, int np_cap, bool mixed, const D& desc) {
    if (inst_end <= inst_begin) return;
    const size_t n_inst = inst_end - inst_begin;
    const int oct_warps = octet_warps_per_sm(oct_ws_doubles);
    if (oct_warps >= 1) {
      const size_t oct_smem = (size_t)4 * oct_ws_doubles * sizeof(double);
      prof_begin();
      allow_smem((const void*)k_solve_oct<D>, oct_smem);
      const size_t blocks_needed = (n_inst + 3) / 4;
      // the U slab of all resident warps should stay in L2 (126 MB): fewer warps for very long paths
      const size_t per_cta = (size_t)4 * std::max(np_cap, 1) * tg::kOctURow;
      size_t grid = std::min(blocks_needed, (size_t)sm_count * oct_warps);
      const size_t l2_ctas = std::max<size_t>((size_t)sm_count, ((size_t)96 << 20) / (per_cta * sizeof(double)));
      grid = std::min(grid, l2_ctas);
      double* us = slab(grid * per_cta);
      k_solve_oct<D><<<(unsigned)grid, 32, oct_smem, stream>>>(desc, inst_begin, inst_end, oct_ws_doubles, std::max(np_cap, 1), us, skip_thread_eligible);
      TG_CUDA_CHECK(cudaGetLastError());
      prof_end(typeid(D).name(), n_inst);
      if (!mixed) return;
    }
    const int skip = oct_warps >= 1 ? oct_ws_doubles : 0;
    const size_t ws_bytes = (size_t)ws_doubles * sizeof(double);
    const size_t smem = ws_bytes * kSolveWarps;
    const size_t blocks_needed = (n_inst + kSolveWarps - 1) / kSolveWarps;
    prof_begin();
    if (smem <= smem_optin) {
      allow_smem((const void*)k_solve<D>, smem);
      int per_sm = 0;
      TG_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_solve<D>, kSolveWarps * 32, smem));
      if (per_sm < 1) per_sm = 1;
      const size_t grid = std::min(blocks_needed, (size_t)sm_count * per_sm);
      k_solve<D><<<(unsigned)grid, kSolveWarps * 32, smem, stream>>>(desc, inst_begin, inst_end, ws_doubles, nullptr, skip, skip_thread_eligible);
    } else {
      // long paths: per-warp workspace in a global slab (L2 resident), persistent grid
      const size_t grid = std::min(blocks_needed, (size_t)sm_count * 8);
      const size_t need = grid * kSolveWarps * (size_t)ws_doubles;
      if (need > gen_slab_doubles) {
        if (gen_slab) {
          TG_CUDA_CHECK(cudaStreamSynchronize(stream));
          TG_CUDA_CHECK(cudaFree(gen_slab));
        }
        TG_CUDA_CHECK(cudaMalloc(&gen_slab, need * sizeof(double)));
        gen_slab_doubles = need;
      }
      k_solve<D><<<(unsigned)grid, kSolveWarps * 32, 0, stream>>>(desc, inst_begin, inst_end, ws_doubles, gen_slab, skip, skip_thread_eligible);
    }
    TG_CUDA_CHECK(cudaGetLastError());
    prof_end(typeid(D).name(), n_inst);
  }

  // out[0..count) = indices i < n with flags[i] != 0, ascending (order-preserving compaction); *count = how many (device)
  void select_flagged(const uint8_t* flags, int* out, int* count, int n) {
    if (n <= 0) {
      TG_CUDA_CHECK(cudaMemsetAsync(count, 0, sizeof(int), stream));
      return;
    }
    cub::CountingInputIterator<int> idx(0);
    size_t bytes = 0;
    TG_CUDA_CHECK(cub::DeviceSelect::Flagged(nullptr, bytes, idx, flags, out, count, n, stream));
    if (bytes > scan_tmp_bytes) {
      if (scan_tmp) {
        TG_CUDA_CHECK(cudaStreamSynchronize(stream));
        TG_CUDA_CHECK(cudaFree(scan_tmp));
      }
      TG_CUDA_CHECK(cudaMalloc(&scan_tmp, bytes));
      scan_tmp_bytes = bytes;
    }
    TG_CUDA_CHECK(cub::DeviceSelect::Flagged(scan_tmp, bytes, idx, flags, out, count, n, stream));
  }
  // out[0..n] = exclusive prefix sums of in[0..n-1], out[n] = total.  `in` must have n+1 elements.
  void exclusive_scan(int* in, int* out, int n) {
    TG_CUDA_CHECK(cudaMemsetAsync(in + n, 0, sizeof(int), stream));
    size_t bytes = 0;
    TG_CUDA_CHECK(cub::DeviceScan::ExclusiveSum(nullptr, bytes, in, out, n + 1, stream));
    if (bytes > scan_tmp_bytes) {
      if (scan_tmp) {
        TG_CUDA_CHECK(cudaStreamSynchronize(stream));
        TG_CUDA_CHECK(cudaFree(scan_tmp));
      }
      TG_CUDA_CHECK(cudaMalloc(&scan_tmp, bytes));
      scan_tmp_bytes = bytes;
    }
    TG_CUDA_CHECK(cub::DeviceScan::ExclusiveSum(scan_tmp, bytes, in, out, n + 1, stream));
  }
};

#define TG_BACKEND CudaBackend
#include "tg_capi_impl.hpp"

// ---- CUDA-only entry points (declared in include/tg_b200.h) ------------------------------------------------------------
extern "C" {

int tg_set_profiling(tg_ctx* ctx, int on) {
  if (!ctx) return TG_ERR_INVALID;
  ctx->be.profiling = on != 0;
  ctx->profiling = on != 0;  // per-kernel timing runs the whole batch on lane 0
  ctx->be.prof.clear();
  return TG_OK;
}

// Writes up to `cap` entries: names as a '\n'-separated list into `names` (size names_cap), ms / launches / items arrays.
int tg_get_profile(tg_ctx* ctx, int cap, char* names, int names_cap, double* ms, long long* launches, long long* items) {
  if (!ctx || !names || !ms || !launches || !items) return TG_ERR_INVALID;
  int n = 0;
  std::string all;
  for (const auto& kv : ctx->be.prof) {
    if (n >= cap) break;
    all += kv.first;
    all += "\n";
    ms[n] = kv.second.ms;
    launches[n] = kv.second.launches;
    items[n] = kv.second.items;
    ++n;
  }
  if ((int)all.size() + 1 > names_cap) return TG_ERR_INVALID;
  std::memcpy(names, all.c_str(), all.size() + 1);
  return n;
}

double tg_measure_fp64_peak(tg_ctx* ctx, int mode) {
  if (!ctx) return 0.0;
  try {
    return ctx->be.fp64_peak_tflops(mode);
  } catch (...) {
    return 0.0;
  }
}

}  // extern "C"
