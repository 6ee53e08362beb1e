// xarm_lib.cu - kernels + the C ABI of libxarm_b200.so (include/xarm_abi.h).  No torch, no CPU fallback:
// every entry point either runs the sm_100a kernels below or returns an error.
#include <cuda_runtime.h>

#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "xarm_pipeline.cuh"
#include "xarm_heavy.cuh"
#include "xarm_vecnorm.cuh"
#include "xarm_her.cuh"

// ------------------------------------------------------------------------------------------------ kernels
// One thread per env; 128-thread blocks (a warp steps 32 envs in lock-step).
template <class T>
__global__ void __launch_bounds__(128) k_init(KArgs a) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < a.n) body_init<T>(a, i);
}

// ------------------------------------------------------------------------------------------------ step pipeline
// (xarm_pipeline.cuh).  Thread t works on env t, or on env list[t] for t < *list_count when the launch carries a list.
__device__ __forceinline__ int64_t pipe_env(const KArgs& a, int64_t t) {
  if (a.list) return t < (int64_t)*a.list_count ? (int64_t)a.list[t] : -1;
  return t < a.n ? t : -1;
}
// Launches are persistent-style: the grid may be smaller than the work (xarm_step keeps a few block slots per SM free
// for its latency-bound early branch), so every kernel walks its envs with a block-stride loop.  Trip counts are
// uniform per block (the bodies use warp-wide ballots).
// Main-branch launches of a split step (a.work != NULL) claim their chunks from a counter instead: the blocks that land
// on the SMs reserved for the early branch leave at once (PIPE_LEAVE_RESERVED), the others share all the work.
__device__ __forceinline__ bool on_reserved_sm(const KArgs& a) {
  unsigned smid;
  asm("mov.u32 %0, %%smid;" : "=r"(smid));
  return (a.sm_mask[(smid >> 6) & 3u] >> (smid & 63u)) & 1ull;
}
// Guard: if EVERY block of a launch landed on reserved SMs (other streams or processes may fill the rest), nobody would be left
// to do the work.  Every block counts its arrival in a.work[1]; the last one to arrive stays and works wherever it runs.
__device__ __forceinline__ bool pipe_is_last_arrival(const KArgs& a) {
  __shared__ int last_;
  if (threadIdx.x == 0) last_ = atomicAdd(a.work + 1, 1) == (int)gridDim.x - 1;
  __syncthreads();
  return last_ != 0;
}
#define PIPE_LEAVE_RESERVED(a) if ((a).work) { const bool res_ = on_reserved_sm(a); if (!pipe_is_last_arrival(a) && res_) return; }
// next chunk (of `chunk` work items) of this block: static stride or claimed from a.work; uniform over the block
__device__ __forceinline__ int64_t pipe_next(const KArgs& a, int64_t prev, int chunk) {
  if (!a.work) return prev < 0 ? (int64_t)blockIdx.x * chunk : prev + (int64_t)gridDim.x * chunk;
  __shared__ int claim_;
  __syncthreads();   // (everyone has read the previous claim)
  if (threadIdx.x == 0) claim_ = atomicAdd(a.work, 1);
  __syncthreads();
  return (int64_t)claim_ * chunk;
}
// Short lists (the auto-reset tail of an ordinary step: a handful of envs) are SPREAD over the lanes - one env per 32 or 8
// lanes - so that a warp does not serialise the different code paths of its envs (near / far from the lego, resting /
// airborne lego, contact cases); those launches are latency bound and the idle lanes cost nothing.
__device__ __forceinline__ int pipe_spread(const KArgs& a) {
  if (!a.list) return 1;
  const int c = *a.list_count;
  return c <= 64 ? 32 : (c <= 256 ? 8 : 1);
}
// the setup kernel (collision code: many data-dependent paths) also spreads mid-sized lists - the ~3 k finishers a staggered
// step resets - 4 lanes per env: a warp then serialises the paths of 8 envs instead of 32
__device__ __forceinline__ int pipe_spread_setup(const KArgs& a) {
  if (!a.list) return 1;
  const int c = *a.list_count;
  return c <= 64 ? 32 : (c <= 640 ? 8 : (c <= a.spread4_max ? 4 : 1));
}
#define PIPE_FOR_EACH(a, t, i) PIPE_FOR_EACH_SP(a, t, i, pipe_spread(a))
#define PIPE_FOR_EACH_SP(a, t, i, spread_)                                                                       \
  for (int64_t sp_ = (spread_), bound_ = ((a).list ? (int64_t)*(a).list_count : (a).n) * sp_,              \
               base_ = pipe_next(a, -1, blockDim.x);                                                            \
       base_ < bound_; base_ = pipe_next(a, base_, blockDim.x))                                                  \
    if (const int64_t tv_ = base_ + threadIdx.x; true)                                                           \
      if (const int64_t t = tv_ / sp_; true)                                                                     \
        if (const int64_t i = (tv_ % sp_ == 0) ? pipe_env(a, t) : -1; true)

// development timeline: thread 0 of every block stamps the launch's slot (min of starts, max of ends)
__device__ __forceinline__ void tl_mark(const KArgs& a, int end) {
  if (a.tl && threadIdx.x == 0 && a.tl_slot >= 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    if (end) atomicMax(a.tl + 2 * a.tl_slot + 1, t); else atomicMin(a.tl + 2 * a.tl_slot, t);
  }
}
// warp-aggregated append of the flagged lanes' env ids to a list
__device__ __forceinline__ void list_append(bool flag, int64_t i, int* list, int* count, int dir = 1) {
  const unsigned m = __ballot_sync(0xffffffffu, flag);
  if (!m) return;
  const int lane = threadIdx.x & 31, leader = __ffs(m) - 1;
  int base = 0;
  if (lane == leader) base = atomicAdd(count, __popc(m));
  base = __shfl_sync(0xffffffffu, base, leader);
  if (flag) list[dir * (base + __popc(m & ((1u << lane) - 1u)))] = (int)i;
}

template <class T>
__global__ void __launch_bounds__(128) k_pipe_action(KArgs a) {
  PIPE_LEAVE_RESERVED(a)
  tl_mark(a, 0);
  PIPE_FOR_EACH(a, t, i) { if (i >= 0) pipe_action<T>(a, i); }
  tl_mark(a, 1);
}
#ifndef XARM_SETUP_MINB
#define XARM_SETUP_MINB 4
#endif
#ifndef XARM_LIGHT_MINB
#define XARM_LIGHT_MINB 4
#endif
// (A 255-register form of this kernel with the link passes unrolled - every array in registers - was measured for the short
// lists of the auto-reset tail: 70 us per launch against 71 us.  ncu on tail-sized launches: no_instruction is 8 of 12.5 stall
// cycles per issue - a lone warp per SM is bound by instruction fetch and branch bubbles of ~130 KB of once-executed code, not
// by the thread-local arrays.  profiles/r2c_*.)
#define XARM_LAT_MAX 16384   // (the ~7 k candidates of a staggered step take the latency forms too: one block of 128 per SM)
template <class T, bool LAT>
__device__ __forceinline__ void setup_body(const KArgs& a, int sub, int* heavy_count) {
  if (a.list && a.light_dual && LAT != (*a.list_count <= XARM_LAT_MAX)) return;
  if (a.tl && blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(a.tl + 5 * XARM_TL_SLOTS + 3 + a.tl_branch, (unsigned long long)(a.list ? *a.list_count : a.n));
  PIPE_LEAVE_RESERVED(a)
  tl_mark(a, 0);
  PIPE_FOR_EACH_SP(a, t, i, pipe_spread_setup(a)) {
    const bool heavy = i >= 0 && pipe_setup<T, LAT>(a, i, sub);
    list_append(heavy, i, a.heavy_list, heavy_count, a.heavy_dir);
  }
  tl_mark(a, 1);
}
// Register budget of the setup kernel.  Single-island-pair tasks (Reach, PickAndPlace with one lego): the FLAT form - link passes
// unrolled, f[] / I[] / S[] / M[] in registers, 255 registers, 2 blocks per SM.  The rolled 128-register form (4 blocks per SM) keeps a
// 2.1 KB thread-local stack frame per thread: 165 MB over the resident threads of a full-size launch, more than the L2 holds - 270 MB
// read + 330 MB written per launch (profiles/r2o_traffic_summary.txt), and that stream evicts the early branch's working set from the
// L2: with the FLAT form the main branch's launch takes the same 277 us at half the warps and the early branch's kernels run 8-13 %
// faster next to it (setup 71 -> 62 us, fused heavy 171 -> 159 us; 4.93 -> 5.07 M env-steps/s).  XARM_SETUP_FLAT=0/1 overrides (A/B).
template <class T>
constexpr bool setup_flat() {
#ifdef XARM_SETUP_FLAT
  return XARM_SETUP_FLAT != 0;
#else
  return task_single_island_pair<T>();
#endif
}
template <class T>
__global__ void __launch_bounds__(128, setup_flat<T>() ? 2 : XARM_SETUP_MINB) k_pipe_setup(KArgs a, int sub, int* heavy_count) { setup_body<T, setup_flat<T>()>(a, sub, heavy_count); }

// Two register budgets of the same light kernel.  LAT = false: 128 registers (some spills), 4 blocks per SM - the
// throughput form (main branch, reset waves).  LAT = true: 160 registers (no hot spills: 86 instead of 121 us per warp) -
// the latency form for SHORT lists (the auto-reset tail of an ordinary step).  A list launch carries both; each looks at
// the list size and one of them leaves at once (the register budget is fixed per launch, the list size is not known on
// the host).
#define XARM_LIGHT_LAT_MAX XARM_LAT_MAX
template <class T, bool LAT, bool REGROWS = false>
__device__ __forceinline__ void light_body(const KArgs& a) {
  extern __shared__ float light_mrows[];  // [XARM_MROW_WORDS][128]: the manifold rows of this block's envs
  PIPE_LEAVE_RESERVED(a)
  if (a.list && a.light_dual && LAT != (*a.list_count <= XARM_LIGHT_LAT_MAX)) return;
  tl_mark(a, 0);
  if constexpr (REGROWS) {   // tail form: the manifold rows in registers too (static indices after unrolling)
    PIPE_FOR_EACH(a, t, i) { if (i >= 0) { float mr[XARM_MROW_WORDS]; pipe_light<T>(a, i, mr, 1); } }
  } else {
    PIPE_FOR_EACH(a, t, i) { if (i >= 0) pipe_light<T>(a, i, light_mrows + threadIdx.x, 128); }
  }
  tl_mark(a, 1);
}
template <class T>
__global__ void __launch_bounds__(128, 4) k_pipe_light(KArgs a) { light_body<T, false>(a); }
#ifndef XARM_LIGHT_LAT_REGS
#define XARM_LIGHT_LAT_REGS 160
#endif
template <class T>
__global__ void __maxnreg__(XARM_LIGHT_LAT_REGS) k_pipe_light_lat(KArgs a) { light_body<T, true>(a); }
// tail form: 255 registers, no shared memory - for the short lists of the early branch (one warp per SM)
template <class T>
__global__ void __maxnreg__(255) k_pipe_light_tail(KArgs a) { light_body<T, true, true>(a); }
// Heavy envs of this substep.  One warp per block, a few blocks per SM at most: the contact rows of the generic solver
// (Contacts<T>, ~6 KB per env) live in SHARED memory, one record per lane at an odd word stride (bank-conflict free).
// In thread-local memory the 50 sweeps stream every row from L2 again (160 KB per warp per sweep) and one heavy warp
// needs > 1 ms per substep; the launch is persistent (grid = #SMs) and runs next to k_pipe_light.
template <class T>
constexpr bool task_has_heavy_rows() { return T::NARM == 1 && T::NOBJ == 1 && !T::HAS_DOOR && T::MD::N == 9; }
template <class T>
constexpr int heavy_stride_words() { return (int)((sizeof(Contacts<T>) + 3) / 4) | 1; }
template <class T>
constexpr size_t heavy_smem_bytes_of() {
  if constexpr (task_has_heavy_rows<T>()) return HeavyLayout<T>::BYTES;
  else return (size_t)heavy_stride_words<T>() * 32 * sizeof(float);
}
template <class T>
__global__ void __launch_bounds__(32) k_pipe_heavy(KArgs a, int sub, const int* heavy_count) {
  extern __shared__ float4 heavy_smem4[];
  float* heavy_smem = reinterpret_cast<float*>(heavy_smem4);
  const int count = *heavy_count;
  for (int t = blockIdx.x * 32 + threadIdx.x; t < count; t += gridDim.x * 32) {
    const int64_t i = a.heavy_list[a.heavy_dir * t];
    if constexpr (task_has_heavy_rows<T>()) {  // solver rows as float4 records (xarm_heavy.cuh)
      Env<T> e;
      env_load<T>(e, a.state, a.n, i);
      heavy_substep<T>(e, T::DAMP_EACH || sub == 0, sub == T::NSUB - 1, heavy_smem, threadIdx.x);
      env_store<T>(e, a.state, a.n, i);
    } else {                                   // the generic record, one per lane at an odd word stride
      Contacts<T>& C = *reinterpret_cast<Contacts<T>*>(heavy_smem + (size_t)threadIdx.x * heavy_stride_words<T>());
      pipe_heavy<T>(a, i, sub, C);
    }
  }
}
// Tasks whose generic contact record does not fit shared memory 32 lanes wide (two arms, several objects, the door: ~9 KB per
// env): the heavy envs of the substep - the minority there - keep the record in thread-local memory, one thread per env.
template <class T>
constexpr bool heavy_record_fits_smem() { return heavy_smem_bytes_of<T>() <= 160 * 1024; }
// Two forms, both launched, the list size decides on the device (as for the other heavy kernels):
//  RECORD IN SHARED MEMORY (lists <= XARM_HEAVY_REC_MAX): 8 envs per block, one lane each, the env's record packed in shared
//    memory (3 blocks = 24 envs per SM).  Thread-local memory is interleaved over the 32 lanes of a warp, so a warp's records
//    occupy 32 x 9 KB of cache whatever the number of active lanes: the 50 sweeps then stream every row from L2 (2-3 ms per
//    substep measured for ANY list size, one env per thread as well as one env per 8 lanes).
//  DENSE, thread-local (longer lists: the reset wave of a batch whose episodes run in phase): one env per thread.
#define XARM_HEAVY_REC_MAX 12288
#define XARM_HEAVY_REC_LANES 8
template <class T>
constexpr size_t heavy_rec_smem_bytes() { return (size_t)heavy_stride_words<T>() * XARM_HEAVY_REC_LANES * sizeof(float); }
template <class T>
__global__ void __launch_bounds__(XARM_HEAVY_REC_LANES) k_pipe_heavy_rec(KArgs a, int sub, const int* heavy_count) {
  extern __shared__ float4 heavy_smem4[];
  float* heavy_smem = reinterpret_cast<float*>(heavy_smem4);
  const int count = *heavy_count;
  if (count > XARM_HEAVY_REC_MAX) return;
  if (a.tl && blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(a.tl + 5 * XARM_TL_SLOTS + a.tl_branch, (unsigned long long)count);
  PIPE_LEAVE_RESERVED(a)
  bool any = false;
  for (int64_t base = pipe_next(a, -1, XARM_HEAVY_REC_LANES); base < count; base = pipe_next(a, base, XARM_HEAVY_REC_LANES)) {
    if (!any) { tl_mark(a, 0); any = true; }
    const int64_t t = base + threadIdx.x;
    if (t < count) {
      Contacts<T>& C = *reinterpret_cast<Contacts<T>*>(heavy_smem + (size_t)threadIdx.x * heavy_stride_words<T>());
      pipe_heavy<T>(a, a.heavy_list[a.heavy_dir * t], sub, C);
    }
  }
  if (any) tl_mark(a, 1);
}
template <class T>
__global__ void __launch_bounds__(64) k_pipe_heavy_local(KArgs a, int sub, const int* heavy_count) {
  const int count = *heavy_count;
  if (count <= XARM_HEAVY_REC_MAX) return;
  if (a.tl && blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(a.tl + 5 * XARM_TL_SLOTS + a.tl_branch, (unsigned long long)count);
  PIPE_LEAVE_RESERVED(a)
  bool any = false;
  for (int64_t base = pipe_next(a, -1, 64); base < count; base = pipe_next(a, base, 64)) {
    if (!any) { tl_mark(a, 0); any = true; }
    const int64_t t = base + threadIdx.x;
    if (t < count) { Contacts<T> C; pipe_heavy<T>(a, a.heavy_list[a.heavy_dir * t], sub, C); }
  }
  if (any) tl_mark(a, 1);
}
// Cooperative heavy path (tasks with HeavyRec: PickAndPlace).  k_heavy_rows: one thread per heavy env, generic setup ->
// record in global memory.  k_heavy_solve: 16 lanes per env, 8 envs per 128-thread block, records staged in shared
// memory (3 blocks = 24 envs = 12 warps per SM instead of the single warp of the thread-per-env form).
#define XARM_HEAVY_ENVS_PER_BLOCK 8
// Lists of heavy envs up to this size take the fused kernel (latency form: 5 warps per SM), longer ones - the reset wave of a batch
// whose episodes run in phase, the steps after it - k_heavy_rows + k_heavy_solve2 (throughput forms).  Both are launched; the
// list size, known on the device only, decides which one works.
#define XARM_FUSED_MAX 4096
template <class T>
__global__ void __launch_bounds__(64) k_heavy_rows(KArgs a, int sub, const int* heavy_count, float* hrec) {
  if constexpr (task_has_heavy_rows<T>()) {
    if (a.heavy_dual == 2 && *heavy_count <= XARM_FUSED_MAX) return;
    if (a.tl && blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(a.tl + 5 * XARM_TL_SLOTS + a.tl_branch, (unsigned long long)*heavy_count);
    PIPE_LEAVE_RESERVED(a)
    const int count = *heavy_count;
    if (a.tl && a.tl_slot >= 0 && threadIdx.x == 0) a.tl[2 * XARM_TL_SLOTS + a.tl_slot] = (unsigned long long)count;
    // Short lists (the auto-reset tail, the candidates of the early branch) are spread over the lanes: with one env per
    // 8 or 32 lanes the warp no longer serialises the different collision cases (face / edge contacts, 1..4 manifolds) of
    // 32 envs - the kernel is latency bound there and the idle lanes cost nothing.
    const int spread = count <= 256 ? 32 : (count <= 1024 ? 8 : 1);
    const int64_t vcount = (int64_t)count * spread;
    bool any = false;
    for (int64_t base = pipe_next(a, -1, blockDim.x); base < vcount; base = pipe_next(a, base, blockDim.x)) {
      if (!any) { tl_mark(a, 0); any = true; }
      const int64_t g = base + threadIdx.x;
      if (g < vcount && g % spread == 0) heavy_rows_body<T>(a, (int)(g / spread), sub, hrec);
    }
    if (any) tl_mark(a, 1);
  }
}
template <class T>
__global__ void __launch_bounds__(128) k_heavy_solve(KArgs a, const int* heavy_count, const float* hrec) {
  if constexpr (task_has_heavy_rows<T>()) {
    extern __shared__ float4 heavy_smem4[];
    constexpr int SLOT = HeavyRec<T>::WORDS + T::MAXC * 3;
    const int g = threadIdx.x >> 4, l = threadIdx.x & 15;
    float* srec = reinterpret_cast<float*>(heavy_smem4) + (size_t)g * SLOT;
    const int count = *heavy_count;
    for (int base = blockIdx.x * XARM_HEAVY_ENVS_PER_BLOCK; base < count; base += gridDim.x * XARM_HEAVY_ENVS_PER_BLOCK)
      heavy_solve_body<T>(a, base + g, base + g < count, hrec, srec, l);
  }
}
// Low-latency form (heavy_solve_dela: impulse-space joint loop, ~4x shorter dependency chain per row): one warp = 2 envs
// per block, 15.8 KB of shared memory per env (the Delassus matrix of up to 58 rows), 7 blocks per SM.
#ifndef XARM_HEAVY2_ENVS_PER_BLOCK
#define XARM_HEAVY2_ENVS_PER_BLOCK 2
#endif
#define XARM_HEAVY2_BLOCKS_PER_SM 7
template <class T>
__global__ void __launch_bounds__(16 * XARM_HEAVY2_ENVS_PER_BLOCK) k_heavy_solve2(KArgs a, const int* heavy_count, const float* hrec) {
  if constexpr (task_has_heavy_rows<T>()) {
    extern __shared__ float4 heavy_smem4[];
    const int g = threadIdx.x >> 4, l = threadIdx.x & 15;
    float* sm = reinterpret_cast<float*>(heavy_smem4) + (size_t)g * DelaLayout<T>::SLOT;
    if (a.heavy_dual == 2 && *heavy_count <= XARM_FUSED_MAX) return;
    PIPE_LEAVE_RESERVED(a)
    const int count = *heavy_count;
    bool any = false;
    for (int64_t base = pipe_next(a, -1, XARM_HEAVY2_ENVS_PER_BLOCK); base < count; base = pipe_next(a, base, XARM_HEAVY2_ENVS_PER_BLOCK)) {
      if (!any) { tl_mark(a, 0); any = true; }
      heavy_solve2_body<T>(a, (int)base + g, base + g < count, hrec, sm, l);
    }
    if (any) tl_mark(a, 1);
  }
}
// Fused form (heavy_fused_body): collision + rows + joint loop in one launch, the record never leaves shared memory.
#define XARM_FUSED_BLOCKS_PER_SM 5
template <class T>
__global__ void __launch_bounds__(32) k_heavy_fused(KArgs a, int sub, const int* heavy_count) {
  if constexpr (task_has_heavy_rows<T>()) {
    extern __shared__ float4 heavy_smem4[];
    const int g = threadIdx.x >> 4, l = threadIdx.x & 15;
    float* sm = reinterpret_cast<float*>(heavy_smem4) + (size_t)g * FusedLayout<T>::SLOT;
    if (a.heavy_dual == 1 && *heavy_count > XARM_FUSED_MAX) return;
    if (a.tl && blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(a.tl + 5 * XARM_TL_SLOTS + a.tl_branch, (unsigned long long)*heavy_count);
    PIPE_LEAVE_RESERVED(a)
    const int count = *heavy_count;
    if (a.tl && a.tl_slot >= 0 && threadIdx.x == 0) a.tl[2 * XARM_TL_SLOTS + a.tl_slot] = (unsigned long long)count;
    bool any = false;
    for (int64_t base = pipe_next(a, -1, 2); base < count; base = pipe_next(a, base, 2)) {
      if (!any) { tl_mark(a, 0); any = true; }
      heavy_fused_body<T>(a, (int)base + g, base + g < count, sub, sm, l);
    }
    if (any) tl_mark(a, 1);
  }
}
// tasks without a light form (two arms / door): every env takes the generic substep (rows in thread-local memory)
template <class T>
__global__ void __launch_bounds__(128) k_pipe_heavy_all(KArgs a, int sub) {
  tl_mark(a, 0);
  PIPE_FOR_EACH(a, t, i) {
    if (i >= 0) { Contacts<T> C; pipe_heavy<T>(a, i, sub, C); }
  }
  tl_mark(a, 1);
}
template <class T>
__global__ void __launch_bounds__(128) k_pipe_finish(KArgs a) {
  PIPE_LEAVE_RESERVED(a)
  tl_mark(a, 0);
  PIPE_FOR_EACH(a, t, i) {
    StepStats st = {0.f, 0.f, 0.f, 0.f, 0.f};
    const bool fin = i >= 0 && pipe_finish<T>(a, i, st);
    list_append(fin, i, a.reset_list, a.reset_count);
    // episode statistics (K8): warp-aggregate, one atomic per warp and counter
    unsigned any = __ballot_sync(0xffffffffu, st.eps != 0.f || st.div != 0.f);
    if (any) {
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) {
        st.eps += __shfl_down_sync(0xffffffffu, st.eps, off);
        st.ret += __shfl_down_sync(0xffffffffu, st.ret, off);
        st.len += __shfl_down_sync(0xffffffffu, st.len, off);
        st.suc += __shfl_down_sync(0xffffffffu, st.suc, off);
        st.div += __shfl_down_sync(0xffffffffu, st.div, off);
      }
      if ((threadIdx.x & 31) == 0) {
        atomicAdd(&a.stats[0], (double)st.eps); atomicAdd(&a.stats[1], (double)st.ret); atomicAdd(&a.stats[2], (double)st.len);
        atomicAdd(&a.stats[3], (double)st.suc);
        if (st.div != 0.f) atomicAdd(&a.stats[4], (double)st.div);
      }
    }
  }
  tl_mark(a, 1);
}
template <class T>
__global__ void __launch_bounds__(128) k_pipe_reset_stage(KArgs a, int stage, int clear_return) {
  tl_mark(a, 0);
  PIPE_FOR_EACH(a, t, i) { if (i >= 0) pipe_reset_stage<T>(a, i, stage, clear_return != 0); }
  tl_mark(a, 1);
}
// which SM ids exist on this device (ids need not be contiguous): every block reports its %smid and lingers a little so
// that the grid spreads over all SMs
__global__ void k_probe_smid(unsigned* seen) {
  unsigned smid;
  asm("mov.u32 %0, %%smid;" : "=r"(smid));
  if (threadIdx.x == 0 && smid < 256u) atomicOr(&seen[smid >> 5], 1u << (smid & 31u));
  const long long t0 = clock64();
  while (clock64() - t0 < 20000) {}
}
// zero the per-launch counters of one env step (heavy lists of every pass and substep, the reset / branch lists)
__global__ void k_pipe_begin(KArgs a, int* counters, int n_counters) {
  for (int t = threadIdx.x; t < n_counters; t += blockDim.x) counters[t] = 0;
  if (a.tl) {  // fold the previous step's stamps into the per-launch accumulators, then re-arm
    for (int t = threadIdx.x; t < XARM_TL_SLOTS; t += blockDim.x) {
      const unsigned long long t0 = a.tl[2 * t], t1 = a.tl[2 * t + 1];
      if (t0 != ~0ull && t1 >= t0) { a.tl[3 * XARM_TL_SLOTS + t] += t1 - t0; a.tl[4 * XARM_TL_SLOTS + t] += 1ull; }
      a.tl[2 * t] = ~0ull; a.tl[2 * t + 1] = 0ull;
    }
  }
}
// envs that may finish in the coming step -> early list, the others -> main list (order-preserving per warp)
template <class T>
__global__ void __launch_bounds__(128) k_pipe_split(KArgs a, int* list_e, int* count_e, int* list_m, int* count_m) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool valid = i < a.n;
  const bool early = valid && pipe_may_finish<T>(a, i);
  list_append(early, i, list_e, count_e);
  list_append(valid && !early, i, list_m, count_m);
}
// list = the envs selected by a mask (xarm_reset)
__global__ void __launch_bounds__(128) k_mask_to_list(KArgs a, const uint8_t* mask) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool on = i < a.n && (!mask || mask[i]);
  list_append(on, i, a.reset_list, a.reset_count);
}

template <class T>
__global__ void __launch_bounds__(128) k_obs(KArgs a) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < a.n) body_obs<T>(a, i);
}

// Env.compute_reward on batches (HER relabelling)
__global__ void k_compute_reward(int task, int reward_type, int num_obj, int G, const float* __restrict__ ag,
                                 const float* __restrict__ dg, int64_t n, float* __restrict__ out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float a[9], d[9];
  for (int k = 0; k < G; k++) { a[k] = ag[i * G + k]; d[k] = dg[i * G + k]; }
  out[i] = reward_stateless(task, reward_type, num_obj, task_threshold(task), a, d, G);
}

// numpy-facing path: rows of the terminal slab whose env finished in this step -> compact rows + env ids (only those cross PCIe)
__global__ void __launch_bounds__(256) k_gather_terminal(const uint8_t* __restrict__ done, const float* __restrict__ term, int64_t n, int W,
                                                         float* __restrict__ rows, int* __restrict__ ids, int* __restrict__ count) {
  const int lane = threadIdx.x & 31;
  for (int64_t base = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) - lane; base < n; base += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = base + lane;
    const bool on = i < n && done[i];
    const unsigned m = __ballot_sync(0xffffffffu, on);
    if (!m) continue;
    int slot0 = 0;
    if (lane == 0) slot0 = atomicAdd(count, __popc(m));
    slot0 = __shfl_sync(0xffffffffu, slot0, 0);
    if (on) ids[slot0 + __popc(m & ((1u << lane) - 1u))] = (int)i;
    // the warp copies its finished rows together: lanes over the words of a row
    for (unsigned mm = m; mm; mm &= mm - 1) {
      const int src_lane = __ffs(mm) - 1;
      const int slot = slot0 + __popc(m & ((1u << src_lane) - 1u));
      const float* src = term + (base + src_lane) * W;
      for (int w = lane; w < W; w += 32) rows[(int64_t)slot * W + w] = src[w];
    }
  }
}

// ------------------------------------------------------------------------------------------------ host side
static thread_local std::string g_err;
static std::atomic<int64_t> g_launches{0};

static int fail(int code, const std::string& msg) { g_err = msg; return code; }
#define CUDA_TRY(x) do { cudaError_t _e = (x); if (_e != cudaSuccess) return fail(XARM_E_CUDA, std::string(#x) + ": " + cudaGetErrorString(_e)); } while (0)

// fork/join plumbing of the pipeline: a side stream for the heavy kernels and a pool of dependency events
struct PipeCtx {
  cudaStream_t side = nullptr;          // heavy kernels of the main branch
  cudaStream_t e_main = nullptr, e_side = nullptr, e_side2 = nullptr;  // the early branch (high priority): envs that may finish this step
  int *list_e = nullptr, *list_m = nullptr, *reset_list_e = nullptr;   // [N] each
  int *count_e = nullptr, *count_m = nullptr, *reset_count_e = nullptr, *counters = nullptr;
  int n_counters = 0;
  bool split = true;                    // XARM_NO_SPLIT=1: one branch (development A/B)
  bool dela = true;                     // XARM_HEAVY_SOLVER=coop: the velocity-form cooperative solver (A/B, parity tests)
  int64_t max_blocks = 4 * 148 - 32;
  unsigned heavy_grid = 148;  // persistent heavy kernels: a few blocks per SM (set from the device in xarm_create)
  float* hrec = nullptr;      // [N][HeavyRec::WORDS] records of the cooperative heavy solver
  std::vector<cudaEvent_t> ev;
  size_t next_ev = 0;
  cudaEvent_t next() {
    if (next_ev == ev.size()) { cudaEvent_t e; cudaEventCreateWithFlags(&e, cudaEventDisableTiming); ev.push_back(e); }
    return ev[next_ev++];
  }
  // XARM_TRACE_STAGES=1 (development): timing events around every pipeline kernel of a non-captured step, summed by
  // kernel and by main pass / auto-reset tail, printed to stderr by xarm_step after a device synchronise
  bool trace = false;
  struct Mark { const char* name; int tail; cudaEvent_t a, b; };
  std::vector<Mark> marks;
  int cur_tail = 0;
  // XARM_TIMELINE=1 (development): %globaltimer stamps of every pipeline launch of the last step, also inside a graph
  bool timeline = false;
  unsigned long long* tl_dev = nullptr;
  std::vector<std::string> tl_names;
  int cur_slot = -1;
  char cur_branch = 'M';
  // SM partition (XARM_RESERVE_SMS, default 32; 0 = off): while `dyn` is set (main branch of a split step) every launch
  // gets a work counter and the mask of the SMs it must leave to the early branch
  bool dyn = false;
  int rec_bps = 3;          // XARM_REC_BPS: resident blocks per SM of k_pipe_heavy_rec (8 envs each; their thread-local frames compete with the records for the L1 / shared-memory array)
  bool e_swap = true;       // XARM_E_SWAP=0: the early branch's heavy kernels on the side stream, its light kernel on the chain's stream (round 2 start)
  bool light_multi = true;  // XARM_LIGHT_MULTI=0: tasks with several islands (two arms, door, several objects) take the generic substep for every env (round 1)
  bool light_dual = true;   // XARM_LIGHT_DUAL=0: one light kernel form everywhere
  bool main_wait = true;    // XARM_MAIN_WAIT=0: the main branch starts together with the early branch
  bool fused = true;        // XARM_HEAVY_FUSED=0: k_heavy_rows + k_heavy_solve2 (round 1) instead of k_heavy_fused
  bool main_fork = true;    // XARM_MAIN_FORK=0: the main branch of a split step keeps one stream (heavy kernels, then light)
  bool tail_lat = true;     // XARM_TAIL_LAT=0: the early branch keeps the round-1 kernels (128-register setup, 160-register light with shared-memory rows)
  bool light_main_lat = true;    // XARM_LIGHT_MAIN_LAT=0: the partitioned main branch keeps the 128-register form (A/B)
  int setup_bps = 0;   // resident blocks per SM of the partitioned main branch's setup kernel (XARM_SETUP_BPS; 0 = what the kernel's register budget allows)
  int reserve_sms = 0, n_work = 0, next_work = 0;
  int* work_base = nullptr;
  unsigned long long sm_mask[4] = {0, 0, 0, 0};
  KArgs tl(const KArgs& a) {
    KArgs b = a;
    b.tl = timeline ? tl_dev : nullptr; b.tl_slot = cur_slot;
    b.tl_branch = cur_branch == 'M' ? 0 : (cur_branch == 'E' ? 1 : 2);
    b.work = nullptr;
    if (dyn && reserve_sms > 0 && next_work < n_work) {
      b.work = work_base + next_work; next_work += 2;   // [0] chunk claims, [1] block arrivals
      for (int k = 0; k < 4; k++) b.sm_mask[k] = sm_mask[k];
    }
    return b;
  }
  void begin(const char* name, cudaStream_t s) {
    if (timeline) {
      cur_slot = (int)tl_names.size() < XARM_TL_SLOTS ? (int)tl_names.size() : -1;
      if (cur_slot >= 0) tl_names.push_back(std::string(1, cur_branch) + " " + name);
    }
    if (!trace) return;
    Mark m; m.name = name; m.tail = cur_tail;
    cudaEventCreate(&m.a); cudaEventCreate(&m.b);
    cudaEventRecord(m.a, s);
    marks.push_back(m);
  }
  void end(cudaStream_t s) { if (trace) cudaEventRecord(marks.back().b, s); }
  void report() {
    if (!trace || marks.empty()) return;
    cudaDeviceSynchronize();
    struct Acc { double ms = 0; int n = 0; };
    std::vector<std::pair<std::string, Acc>> acc;
    for (Mark& m : marks) {
      float ms = 0.f;
      cudaEventElapsedTime(&ms, m.a, m.b);
      std::string key = std::string(m.tail ? "tail " : "main ") + m.name;
      size_t k = 0;
      for (; k < acc.size(); k++) if (acc[k].first == key) break;
      if (k == acc.size()) acc.push_back({key, Acc()});
      acc[k].second.ms += ms; acc[k].second.n++;
      cudaEventDestroy(m.a); cudaEventDestroy(m.b);
    }
    if (counters) {
      int cnt[4] = {0, 0, 0, 0};
      cudaMemcpy(cnt, counters + (n_counters - 4), sizeof(cnt), cudaMemcpyDeviceToHost);
      fprintf(stderr, "[xarm lists] late-tail finishers %d | early finishers %d | early branch %d envs | main branch %d envs\n", cnt[0], cnt[1], cnt[2], cnt[3]);
    }
    fprintf(stderr, "[xarm stages]");
    for (auto& kv : acc) fprintf(stderr, " %s: %d x %.1f us = %.2f ms |", kv.first.c_str(), kv.second.n, 1e3 * kv.second.ms / kv.second.n, kv.second.ms);
    fprintf(stderr, "\n");
    marks.clear();
  }
};

struct Ops {
  void (*init)(const KArgs&, cudaStream_t);
  void (*step)(PipeCtx&, const KArgs&, cudaStream_t);
  void (*reset)(PipeCtx&, const KArgs&, const uint8_t*, cudaStream_t);
  void (*obs)(const KArgs&, cudaStream_t);
  int (*prepare)();
  size_t (*hrec_words)();
  int A, O, G, S, scratch_words;
};

template <class T>
struct OpsT {
  static constexpr bool HAS_LIGHT = task_has_light<T>();  // tasks whose envs can take the light solver form
  static dim3 grid(int64_t n) { return dim3((unsigned)((n + 127) / 128)); }
  // persistent-style grid of the pipeline kernels: at most c.max_blocks (4 blocks of 128 threads x 128 registers fit an SM;
  // a few slots per SM stay free so that the early branch never waits for a block slot)
  static dim3 pgrid(const PipeCtx& c, int64_t n) { const int64_t b = (n + 127) / 128; return dim3((unsigned)(b < c.max_blocks ? b : c.max_blocks)); }
  static size_t heavy_smem_bytes() { return heavy_smem_bytes_of<T>(); }
  static size_t coop_smem_bytes() {
    if constexpr (task_has_heavy_rows<T>()) return (size_t)XARM_HEAVY_ENVS_PER_BLOCK * (HeavyRec<T>::WORDS + T::MAXC * 3) * sizeof(float);
    else return 0;
  }
  static size_t hrec_words() {  // per env: record of the cooperative heavy solver (0: task has none)
    if constexpr (task_has_heavy_rows<T>()) return HeavyRec<T>::WORDS; else return 0;
  }
  static size_t fused_smem_bytes() {
    if constexpr (task_has_heavy_rows<T>()) return (size_t)2 * FusedLayout<T>::SLOT * sizeof(float);
    else return 0;
  }
  static size_t dela_smem_bytes() {
    if constexpr (task_has_heavy_rows<T>()) return (size_t)XARM_HEAVY2_ENVS_PER_BLOCK * DelaLayout<T>::SLOT * sizeof(float);
    else return 0;
  }
  static int prepare() {  // opt in to the large dynamic shared memory of the heavy kernels
    if constexpr (HAS_LIGHT) {  // 48 KB of manifold rows + the static chunk-claim word of pipe_next
      int rc = (int)cudaFuncSetAttribute(k_pipe_light<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(XARM_MROW_WORDS * 128 * sizeof(float)));
      if (rc) return rc;
      rc = (int)cudaFuncSetAttribute(k_pipe_light_lat<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(XARM_MROW_WORDS * 128 * sizeof(float)));
      if (rc) return rc;
    }
    if constexpr (task_has_heavy_rows<T>()) {
      int rc = (int)cudaFuncSetAttribute(k_heavy_solve<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)coop_smem_bytes());
      if (rc) return rc;
      rc = (int)cudaFuncSetAttribute(k_heavy_solve2<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dela_smem_bytes());
      if (rc) return rc;
      return (int)cudaFuncSetAttribute(k_heavy_fused<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fused_smem_bytes());
    }
    else if constexpr (heavy_record_fits_smem<T>()) return (int)cudaFuncSetAttribute(k_pipe_heavy<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)heavy_smem_bytes());
    else return (int)cudaFuncSetAttribute(k_pipe_heavy_rec<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)heavy_rec_smem_bytes<T>());
  }
  static void init(const KArgs& a, cudaStream_t s) { k_init<T><<<grid(a.n), 128, 0, s>>>(a); g_launches++; }
  static void obs(const KArgs& a, cudaStream_t s) { k_obs<T><<<grid(a.n), 128, 0, s>>>(a); g_launches++; }
  // p.stepSimulation(): NSUB x { setup -> light || heavy }
  // sh != s: the heavy kernels run on sh next to k_pipe_light (fork / join); sh == s: one stream, heavy kernels first.
  // The main branch of a split step uses one stream: its kernels fill the GPU anyway, and with a single kernel in
  // flight at a time the block slots that pgrid() leaves free really stay free for the early branch.
  static void simulate(PipeCtx& c, const KArgs& a, int pass, cudaStream_t s, cudaStream_t sh, cudaStream_t sh2 = nullptr) {
    cudaEvent_t join2_pending = nullptr;
    const bool part = c.dyn && c.reserve_sms > 0;   // partitioned main branch: full grids, dynamic chunks
    const dim3 g = part ? dim3(c.heavy_grid * 4) : pgrid(c, a.n);
    const bool fork_heavy = sh != s;
    const unsigned rows_grid = fork_heavy || part ? c.heavy_grid * 4 : (unsigned)c.max_blocks;
    const unsigned solve_grid = fork_heavy || part ? c.heavy_grid * 3 : (unsigned)(c.max_blocks * 3 / 4);
    for (int sub = 0; sub < T::NSUB; sub++) {
      if (!task_single_island_pair<T>() && !c.light_multi) {   // XARM_LIGHT_MULTI=0 (A/B): every env through the generic substep
        c.begin("heavy_all", s);
        k_pipe_heavy_all<T><<<g, 128, 0, s>>>(c.tl(a), sub); g_launches++;
        c.end(s);
      } else {
        int* hc = a.heavy_count + pass * XARM_MAX_SUBSTEPS + sub;
        c.begin("setup", s);
        k_pipe_setup<T><<<part ? dim3(c.heavy_grid * (c.setup_bps > 0 ? c.setup_bps : (setup_flat<T>() ? 2 : 4))) : g, 128, 0, s>>>(c.tl(a), sub, hc);
        c.end(s);
        // Early branch (XARM_E_SWAP, default on): the fused heavy kernel - the critical path of every substep there - stays on the
        // chain's own stream right behind the setup kernel, and the light kernel takes the side stream.
        const bool swap = c.e_swap && c.cur_branch == 'E' && fork_heavy && task_has_heavy_rows<T>() && c.fused && c.dela;
        const cudaStream_t hs = swap ? s : sh, ls = swap ? sh : s;
        cudaEvent_t join = nullptr, fork = nullptr;
        if (fork_heavy) {
          fork = c.next();
          join = c.next();
          cudaEventRecord(fork, s);
          cudaStreamWaitEvent(sh, fork, 0);
        }
        if constexpr (task_has_heavy_rows<T>()) {
          cudaEvent_t join2 = nullptr;
          if (c.fused && c.dela) {
            c.begin("heavy_fused", hs);
            KArgs af = c.tl(a); af.heavy_dual = 1;
            k_heavy_fused<T><<<solve_grid / 3 * XARM_FUSED_BLOCKS_PER_SM, 32, fused_smem_bytes(), hs>>>(af, sub, hc);
            c.end(hs);
          }
          {
            // the throughput pair: after the fused kernel on the same stream, or - early branch, where every launch on the chain
            // counts - next to it on a third stream
            cudaStream_t sr = hs;
            if (c.fused && c.dela && sh2 && fork_heavy) {
              join2 = c.next();
              cudaStreamWaitEvent(sh2, fork, 0);
              sr = sh2;
            }
            c.begin("heavy_rows", sr);
            KArgs ar = c.tl(a); ar.heavy_dual = (c.fused && c.dela) ? 2 : 0;
            k_heavy_rows<T><<<rows_grid, 64, 0, sr>>>(ar, sub, hc, c.hrec);
            c.end(sr);
            c.begin("heavy_solve", sr);
            KArgs as = c.tl(a); as.heavy_dual = ar.heavy_dual;
            if (c.dela) k_heavy_solve2<T><<<solve_grid / 3 * XARM_HEAVY2_BLOCKS_PER_SM, 16 * XARM_HEAVY2_ENVS_PER_BLOCK, dela_smem_bytes(), sr>>>(as, hc, c.hrec);
            else k_heavy_solve<T><<<solve_grid, 128, coop_smem_bytes(), sr>>>(as, hc, c.hrec);
            c.end(sr);
            g_launches += 2;
            if (join2) { cudaEventRecord(join2, sh2); join2_pending = join2; }
          }
        } else if constexpr (heavy_record_fits_smem<T>()) {
          k_pipe_heavy<T><<<c.heavy_grid, 32, heavy_smem_bytes(), sh>>>(a, sub, hc);
        } else {
          c.begin("heavy_rec", sh);
          k_pipe_heavy_rec<T><<<c.heavy_grid * c.rec_bps, XARM_HEAVY_REC_LANES, heavy_rec_smem_bytes<T>(), sh>>>(c.tl(a), sub, hc);
          c.end(sh);
          c.begin("heavy_local", sh);
          k_pipe_heavy_local<T><<<rows_grid, 64, 0, sh>>>(c.tl(a), sub, hc);
          c.end(sh);
          g_launches++;
        }
        c.begin("light", ls);
        if (c.cur_branch == 'E' && c.light_dual) {   // short list -> latency form, long list -> throughput form
          KArgs al = c.tl(a);
          al.light_dual = 1;
          if (c.tail_lat) k_pipe_light_tail<T><<<XARM_LIGHT_LAT_MAX / 128, 128, 0, ls>>>(al);
          else k_pipe_light_lat<T><<<XARM_LIGHT_LAT_MAX / 128, 128, XARM_MROW_WORDS * 128 * sizeof(float), ls>>>(al);
          k_pipe_light<T><<<g, 128, XARM_MROW_WORDS * 128 * sizeof(float), ls>>>(al);
          g_launches++;
        } else if (part && c.light_main_lat) {   // partitioned main branch: nothing runs next to it on its SMs - 3 blocks of the 160-register form
          k_pipe_light_lat<T><<<c.heavy_grid * 3, 128, XARM_MROW_WORDS * 128 * sizeof(float), ls>>>(c.tl(a));
        } else {
          k_pipe_light<T><<<g, 128, XARM_MROW_WORDS * 128 * sizeof(float), ls>>>(c.tl(a));
        }
        c.end(ls);
        if (fork_heavy) { cudaEventRecord(join, sh); cudaStreamWaitEvent(s, join, 0); }
        if (join2_pending) { cudaStreamWaitEvent(s, join2_pending, 0); join2_pending = nullptr; }
        g_launches += 3;
      }
    }
  }
  // Env.reset() of the envs in r.list (all envs when r.list is NULL)
  static void reset_passes(PipeCtx& c, const KArgs& r, int pass, int clear_return, cudaStream_t s, cudaStream_t sh, cudaStream_t sh2 = nullptr) {
    const dim3 g = pgrid(c, r.n);
    if (reset_has_servo<T>())
      for (int rep = 0; rep < 5; rep++) {
        c.begin("reset_stage", s);
        k_pipe_reset_stage<T><<<g, 128, 0, s>>>(c.tl(r), rep, 0); g_launches++;
        c.end(s);
        simulate(c, r, pass++, s, sh, sh2);
      }
    k_pipe_reset_stage<T><<<g, 128, 0, s>>>(r, XARM_RESET_PLACE, 0); g_launches++;
    simulate(c, r, pass++, s, sh, sh2);
    k_pipe_reset_stage<T><<<g, 128, 0, s>>>(r, XARM_RESET_FINISH, clear_return); g_launches++;
  }
  // one branch of a step: _set_action, stepSimulation, outputs for the envs of a.list (all envs when NULL), then
  // Env.reset() of the envs that finished (VecEnv auto-reset) through the same pipeline
  static void step_branch(PipeCtx& c, const KArgs& a, int pass, bool tail, cudaStream_t s, cudaStream_t sh, cudaEvent_t stepped = nullptr, cudaStream_t sh2 = nullptr) {
    const dim3 g = c.dyn && c.reserve_sms > 0 ? dim3(c.heavy_grid * 4) : pgrid(c, a.n);
    c.cur_tail = tail ? 1 : 0;
    c.begin("action", s);
    k_pipe_action<T><<<g, 128, 0, s>>>(c.tl(a));
    c.end(s);
    simulate(c, a, pass, s, sh, sh2);
    c.begin("finish", s);
    k_pipe_finish<T><<<g, 128, 0, s>>>(c.tl(a));
    c.end(s);
    if (stepped) cudaEventRecord(stepped, s);   // the envs of this branch have made their step (their auto-reset passes follow)
    g_launches += 2;
    if (a.auto_reset && tail) {
      c.cur_tail = 1;
      KArgs r = a;
      r.list = a.reset_list; r.list_count = a.reset_count;
      reset_passes(c, r, pass + 1, 0, s, sh, sh2);
    }
  }
  static void step(PipeCtx& c, const KArgs& a0, cudaStream_t s) {
    static_assert(T::NSUB <= XARM_MAX_SUBSTEPS, "substep counters");
    KArgs a = a0;
    a.list = nullptr; a.list_count = nullptr; a.heavy_dir = 1;
    c.next_ev = 0;
    c.cur_slot = -1; c.tl_names.clear(); k_pipe_begin<<<1, 256, 0, s>>>(c.tl(a), c.counters, c.n_counters); g_launches++;
    c.cur_branch = 'M';
    if (!a.auto_reset || !c.split) {  // one branch: every env, then the auto-reset tail
      step_branch(c, a, 0, true, s, c.side);
      return;
    }
    // Two concurrent branches.  Early: the envs that may finish in this step (time limit, success within reach) and
    // their auto-reset passes - few envs, latency bound (6 more stepSimulation passes), high-priority streams.  Main:
    // all the other envs - throughput bound.  The early branch hides the reset latency behind the main branch.
    k_pipe_split<T><<<grid(a.n), 128, 0, s>>>(a, c.list_e, c.count_e, c.list_m, c.count_m); g_launches++;
    cudaEvent_t fork = c.next(), join = c.next();
    cudaEventRecord(fork, s);
    cudaStreamWaitEvent(c.e_main, fork, 0);
    KArgs e = a;
    e.list = c.list_e; e.list_count = c.count_e;
    e.heavy_list = a.heavy_list + (a.n - 1); e.heavy_dir = -1;
    e.reset_list = c.reset_list_e; e.reset_count = c.reset_count_e;
    c.cur_branch = 'E';
    // The early branch is the critical path of the step: 15 + 90 substeps in sequence.  Its first pass (the step of the ~5 k envs
    // that may finish, ~2 k of them in gripper contact) is throughput work, and next to the main branch it takes 12-13 ms
    // instead of ~5: the main branch (which has 20 ms of slack) starts when that pass is over.
    cudaEvent_t e_stepped = c.main_wait ? c.next() : nullptr;
    step_branch(c, e, XARM_PIPE_PASSES / 2, true, c.e_main, c.e_side, e_stepped, c.e_side2);
    cudaEventRecord(join, c.e_main);
    if (e_stepped) cudaStreamWaitEvent(s, e_stepped, 0);
    KArgs m = a;
    m.list = c.list_m; m.list_count = c.count_m;
    c.cur_branch = 'M';
    c.dyn = true; c.next_work = 0;
    // the main branch's heavy kernels (few envs: latency bound) run on a side stream next to its light kernel
    step_branch(c, m, 0, false, s, c.main_fork ? c.side : s);
    c.dyn = false;
    cudaStreamWaitEvent(s, join, 0);
    // late tail: envs of the main branch that finished although the predictor said no (normally none)
    c.cur_tail = 1; c.cur_branch = 'L';
    KArgs r = a;
    r.list = a.reset_list; r.list_count = a.reset_count;
    reset_passes(c, r, 1, 0, s, c.side);
  }
  static void reset(PipeCtx& c, const KArgs& a0, const uint8_t* mask, cudaStream_t s) {
    KArgs a = a0;
    a.list = nullptr; a.list_count = nullptr; a.heavy_dir = 1;
    c.next_ev = 0;
    c.cur_slot = -1; c.tl_names.clear(); k_pipe_begin<<<1, 256, 0, s>>>(c.tl(a), c.counters, c.n_counters); g_launches++;
    if (mask) {
      k_mask_to_list<<<grid(a.n), 128, 0, s>>>(a, mask); g_launches++;
      a.list = a.reset_list; a.list_count = a.reset_count;
    }
    reset_passes(c, a, 0, 1, s, c.side);
  }
  static Ops make() { Ops o = {init, step, reset, obs, prepare, hrec_words, T::A, T::O, T::G, state_words<T>(), pipe_scratch_words<T>()}; return o; }
};

// XARM_ONLY_TASK=<id> builds a single task (development builds: faster compiles); the shipped library has all five.
#ifdef XARM_ONLY_TASK
#define XARM_HAS_TASK(t) ((t) == XARM_ONLY_TASK)
#else
#define XARM_HAS_TASK(t) 1
#endif
static bool get_ops(int task, int num_obj, Ops* out) {
  switch (task) {
#if XARM_HAS_TASK(0)
    case XARM_TASK_REACH: *out = OpsT<TaskT<XARM_TASK_REACH, 0>>::make(); return true;
#endif
#if XARM_HAS_TASK(1)
    case XARM_TASK_PICK_AND_PLACE:
      if (num_obj == 1) { *out = OpsT<TaskT<XARM_TASK_PICK_AND_PLACE, 1>>::make(); return true; }
      if (num_obj == 2) { *out = OpsT<TaskT<XARM_TASK_PICK_AND_PLACE, 2>>::make(); return true; }
      if (num_obj == 3) { *out = OpsT<TaskT<XARM_TASK_PICK_AND_PLACE, 3>>::make(); return true; }
      return false;
#endif
#if XARM_HAS_TASK(2)
    case XARM_TASK_STACK_TOWER: *out = OpsT<TaskT<XARM_TASK_STACK_TOWER, 3>>::make(); return true;
#endif
#if XARM_HAS_TASK(3)
    case XARM_TASK_PUSH_WITH_DOOR: *out = OpsT<TaskT<XARM_TASK_PUSH_WITH_DOOR, 1>>::make(); return true;
#endif
#if XARM_HAS_TASK(4)
    case XARM_TASK_HANDOVER:
      if (num_obj == 1) { *out = OpsT<TaskT<XARM_TASK_HANDOVER, 1>>::make(); return true; }
      if (num_obj == 2) { *out = OpsT<TaskT<XARM_TASK_HANDOVER, 2>>::make(); return true; }
      return false;
#endif
  }
  return false;
}
static int norm_num_obj(int task, int num_obj) {
  if (task == XARM_TASK_REACH) return 0;
  if (task == XARM_TASK_STACK_TOWER) return 3;
  if (task == XARM_TASK_PUSH_WITH_DOOR) return 1;
  return num_obj;
}

struct XarmHandle {
  XarmConfig cfg;
  Ops ops;
  KArgs k;
  bool bound = false;
  PipeCtx pipe;
  cudaGraphExec_t graph = nullptr;
  cudaStream_t graph_stream = nullptr;
  int64_t graph_launches = 0;
  // device + pinned staging for the *_host entry points
  float* d_io = nullptr;   // actions | obs | ag | dg | reward | success | terminal slab | gathered terminal rows
  uint8_t* d_flags = nullptr;  // done | truncated
  int* d_term = nullptr;   // [1 + N]: number of finished envs of the step | their env ids
  float* h_io = nullptr;
  uint8_t* h_flags = nullptr;
  int* h_term = nullptr;
  cudaGraphExec_t graph_host = nullptr;   // the step on the host-path buffers (captured at the first xarm_step_host)
  cudaStream_t graph_host_stream = nullptr, host_stream = nullptr;
  int64_t graph_host_launches = 0;
  XarmBuffers host_bufs;
  XarmBuffers user_bufs;
  bool user_bound = false;
};

extern "C" {

int xarm_abi_version(void) { return XARM_ABI_VERSION; }
const char* xarm_last_error(void) { return g_err.c_str(); }
int64_t xarm_launch_count(void) { return g_launches.load(); }

int xarm_task_dims(int32_t task, int32_t num_obj, int32_t* act_dim, int32_t* obs_dim, int32_t* goal_dim, int32_t* state_words_out) {
  Ops o;
  if (!get_ops(task, norm_num_obj(task, num_obj), &o)) return fail(XARM_E_INVALID, "xarm_task_dims: unsupported task/num_obj");
  if (act_dim) *act_dim = o.A;
  if (obs_dim) *obs_dim = o.O;
  if (goal_dim) *goal_dim = o.G;
  if (state_words_out) *state_words_out = o.S;
  return XARM_OK;
}

int xarm_destroy(XarmHandle* h);
int xarm_create(const XarmConfig* cfg, XarmHandle** out) {
  if (!cfg || !out) return fail(XARM_E_INVALID, "xarm_create: null argument");
  if (cfg->num_envs <= 0) return fail(XARM_E_INVALID, "xarm_create: num_envs must be positive");
  XarmConfig c = *cfg;
  c.num_obj = norm_num_obj(c.task, c.num_obj);
  Ops ops;
  if (!get_ops(c.task, c.num_obj, &ops)) return fail(XARM_E_INVALID, "xarm_create: unsupported task / num_obj (built: num_obj 1..3 for PickAndPlace, 1..2 for Handover)");
  // reward types the reference defines per task (others raise NotImplementedError there, D6)
  bool ok_reward = c.reward_type == XARM_REWARD_SPARSE || c.reward_type == XARM_REWARD_DENSE ||
                   (c.task == XARM_TASK_PICK_AND_PLACE && c.reward_type == XARM_REWARD_DENSE_O2G) ||
                   (c.task == XARM_TASK_REACH && c.reward_type == XARM_REWARD_DENSE_DIFF);
  if (!ok_reward) return fail(XARM_E_INVALID, "xarm_create: reward_type not implemented for this task");
  // config['use_stand'] [REF xarm_handover.py:391-392]: a 0.07 x 0.06 x 0.01 static box under every goal.  The collider is not
  // built: refuse the config instead of simulating a different world silently.
  if (c.use_stand) return fail(XARM_E_INVALID, "xarm_create: use_stand=True is not implemented (the stand collider under the goal is not simulated)");
  int ndev = 0;
  CUDA_TRY(cudaGetDeviceCount(&ndev));
  if (c.device < 0 || c.device >= ndev) return fail(XARM_E_INVALID, "xarm_create: bad device ordinal");
  CUDA_TRY(cudaSetDevice(c.device));
  XarmHandle* h = new (std::nothrow) XarmHandle();
  if (!h) return fail(XARM_E_NOMEM, "xarm_create: out of host memory");
  h->cfg = c; h->ops = ops;
  memset(&h->k, 0, sizeof(h->k));
  memset(&h->host_bufs, 0, sizeof(XarmBuffers));
  memset(&h->user_bufs, 0, sizeof(XarmBuffers));
  const int64_t n = c.num_envs;
  h->k.n = n; h->k.auto_reset = c.auto_reset;
  h->k.rc.seed = c.seed; h->k.rc.env_index_base = c.env_index_base; h->k.rc.reward_type = c.reward_type;
  h->k.rc.goal_shape = c.goal_shape; h->k.rc.max_episode_steps = c.max_episode_steps; h->k.rc.stagger = c.stagger_phases != 0;
  h->k.rc.init_grasp_rate = c.init_grasp_rate; h->k.rc.goal_ground_rate = c.goal_ground_rate; h->k.rc.same_side_rate = c.same_side_rate;
  // int scratch: reset list | heavy list | form | rng draw | early list | main list | early reset list | counters
  const int n_work = 256;  // work counters of the partitioned main branch (a pair per launch: 2 + 4 launches per substep)
  const int n_counters = XARM_PIPE_COUNTERS + n_work + 4;
  const size_t n_int = (size_t)7 * n + n_counters;
  int prio_lo = 0, prio_hi = 0;
  cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
  cudaError_t e1 = cudaMalloc(&h->k.state, sizeof(float) * ops.S * n);
  cudaError_t e2 = cudaMalloc(&h->k.ep_return, sizeof(float) * n);
  cudaError_t e3 = cudaMalloc(&h->k.need_reset, n);
  cudaError_t e4 = cudaMalloc(&h->k.stats, sizeof(double) * 5);
  cudaError_t e5 = cudaMalloc(&h->k.reset_list, sizeof(int) * n_int);
  cudaError_t e6 = cudaMalloc(&h->k.scratch, sizeof(float) * ops.scratch_words * n);
  cudaError_t e7 = cudaStreamCreateWithFlags(&h->pipe.side, cudaStreamNonBlocking);
  if (e7 == cudaSuccess) e7 = cudaStreamCreateWithPriority(&h->pipe.e_main, cudaStreamNonBlocking, prio_hi);
  if (e7 == cudaSuccess) e7 = cudaStreamCreateWithPriority(&h->pipe.e_side, cudaStreamNonBlocking, prio_hi);
  if (e7 == cudaSuccess) e7 = cudaStreamCreateWithPriority(&h->pipe.e_side2, cudaStreamNonBlocking, prio_hi);
  // every failure from here on leaves through ONE cleanup (xarm_destroy frees whatever was allocated so far)
#define CREATE_FAIL(code, msg) do { std::string m_ = (msg); xarm_destroy(h); cudaGetLastError(); return fail((code), m_); } while (0)
  if (e1 != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess || e4 != cudaSuccess || e5 != cudaSuccess || e6 != cudaSuccess || e7 != cudaSuccess)
    CREATE_FAIL(XARM_E_NOMEM, "xarm_create: cudaMalloc failed");
  {
    int sms = 0;
    cudaError_t ea = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c.device);
    if (ea != cudaSuccess) CREATE_FAIL(XARM_E_CUDA, std::string("cudaDeviceGetAttribute: ") + cudaGetErrorString(ea));
    h->pipe.heavy_grid = (unsigned)(sms > 0 ? sms : 148);
    h->pipe.max_blocks = (int64_t)h->pipe.heavy_grid * 4 - (getenv("XARM_RESERVE_BLOCKS") ? atoi(getenv("XARM_RESERVE_BLOCKS")) : 32);
    cudaError_t ep = (cudaError_t)ops.prepare();
    if (ep != cudaSuccess) CREATE_FAIL(XARM_E_CUDA, std::string("cudaFuncSetAttribute(k_pipe_heavy): ") + cudaGetErrorString(ep));
  }
  if (ops.hrec_words() > 0) {
    if (cudaMalloc(&h->pipe.hrec, sizeof(float) * ops.hrec_words() * n) != cudaSuccess) CREATE_FAIL(XARM_E_NOMEM, "xarm_create: cudaMalloc (heavy records) failed");
  }
  {  // SMs reserved for the early branch of a split step
    // Round 2 (staggered phases: ~2.7 k finishers in EVERY step): the early branch is a latency chain of 105 substeps, ~200 us each
    // when it runs alone (one warp per scheduler, bound by instruction fetch and dependent issue) and ~260 us when its warps share
    // schedulers with the main branch's.  36 reserved SMs: the chain's passes take 3.85 ms instead of 5.5-5.7 next to the main
    // branch (3.0 alone); the main branch, which has ~10 ms of slack, takes 15.6 ms on the other 112 SMs instead of 10.4.
    const int want = getenv("XARM_RESERVE_SMS") ? atoi(getenv("XARM_RESERVE_SMS")) : 36;
    if (want > 0 && ops.hrec_words() > 0) {
      unsigned* d_seen = nullptr;
      unsigned seen[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      if (cudaMalloc(&d_seen, sizeof(seen)) == cudaSuccess) {
        cudaMemset(d_seen, 0, sizeof(seen));
        k_probe_smid<<<h->pipe.heavy_grid * 16, 32>>>(d_seen);
        cudaMemcpy(seen, d_seen, sizeof(seen), cudaMemcpyDeviceToHost);
        cudaFree(d_seen);
        int total = 0, taken = 0;
        for (int id = 0; id < 256; id++) total += (seen[id >> 5] >> (id & 31)) & 1u;
        if (total >= 2 * want)
          for (int id = 0; id < 256 && taken < want; id++)
            if ((seen[id >> 5] >> (id & 31)) & 1u) { h->pipe.sm_mask[id >> 6] |= 1ull << (id & 63); taken++; }
        h->pipe.reserve_sms = taken;
      }
      cudaGetLastError();
    }
  }
  if (getenv("XARM_SETUP_BPS")) h->pipe.setup_bps = atoi(getenv("XARM_SETUP_BPS"));
  if (getenv("XARM_REC_BPS")) h->pipe.rec_bps = atoi(getenv("XARM_REC_BPS")) > 0 ? atoi(getenv("XARM_REC_BPS")) : 3;
  h->pipe.e_swap = !(getenv("XARM_E_SWAP") && atoi(getenv("XARM_E_SWAP")) == 0);
  h->pipe.light_multi = !(getenv("XARM_LIGHT_MULTI") && atoi(getenv("XARM_LIGHT_MULTI")) == 0);
  h->pipe.light_dual = !(getenv("XARM_LIGHT_DUAL") && atoi(getenv("XARM_LIGHT_DUAL")) == 0);
  h->pipe.main_wait = !(getenv("XARM_MAIN_WAIT") && atoi(getenv("XARM_MAIN_WAIT")) == 0);
  h->pipe.fused = !(getenv("XARM_HEAVY_FUSED") && atoi(getenv("XARM_HEAVY_FUSED")) == 0);
  // With the SM partition on, the main branch's heavy kernel must NOT run next to its light kernel: its resident blocks (47 KB of
  // shared memory each) fill the unreserved SMs, the light kernel's blocks then only find room on the reserved SMs - where they
  // leave at once - and the launch's last block does all the work alone (measured: 77 ms per launch).
  h->pipe.main_fork = getenv("XARM_MAIN_FORK") ? atoi(getenv("XARM_MAIN_FORK")) != 0 : h->pipe.reserve_sms == 0;
  h->pipe.tail_lat = !(getenv("XARM_TAIL_LAT") && atoi(getenv("XARM_TAIL_LAT")) == 0);
  h->pipe.light_main_lat = !(getenv("XARM_LIGHT_MAIN_LAT") && atoi(getenv("XARM_LIGHT_MAIN_LAT")) == 0);
  h->pipe.trace = getenv("XARM_TRACE_STAGES") != nullptr;
  h->pipe.timeline = getenv("XARM_TIMELINE") != nullptr;
  if (h->pipe.timeline && cudaMalloc(&h->pipe.tl_dev, sizeof(unsigned long long) * (5 * XARM_TL_SLOTS + 8)) != cudaSuccess) { cudaGetLastError(); h->pipe.timeline = false; }
  if (h->pipe.timeline) cudaMemset(h->pipe.tl_dev, 0, sizeof(unsigned long long) * (5 * XARM_TL_SLOTS + 8));
  h->pipe.split = getenv("XARM_NO_SPLIT") == nullptr;
  h->pipe.dela = !(getenv("XARM_HEAVY_SOLVER") && strcmp(getenv("XARM_HEAVY_SOLVER"), "coop") == 0);
  h->k.heavy_list = h->k.reset_list + n; h->k.form = h->k.reset_list + 2 * n; h->k.rng_draw = h->k.reset_list + 3 * n;
  h->pipe.list_e = h->k.reset_list + 4 * n; h->pipe.list_m = h->k.reset_list + 5 * n; h->pipe.reset_list_e = h->k.reset_list + 6 * n;
  h->pipe.counters = h->k.reset_list + 7 * n; h->pipe.n_counters = n_counters;
  h->k.heavy_count = h->pipe.counters; h->k.reset_count = h->pipe.counters + XARM_PIPE_COUNTERS + n_work;
  h->pipe.work_base = h->pipe.counters + XARM_PIPE_COUNTERS; h->pipe.n_work = n_work;
  h->pipe.reset_count_e = h->k.reset_count + 1; h->pipe.count_e = h->k.reset_count + 2; h->pipe.count_m = h->k.reset_count + 3;
  h->k.heavy_dir = 1;
  h->k.heavy_dual = 0;
  h->k.spread4_max = getenv("XARM_SETUP_SPREAD4") ? atoi(getenv("XARM_SETUP_SPREAD4")) : 0;   // measured neutral at 3.4 k envs (56 vs 62 us in quiet passes, slower next to the main branch): off
  cudaError_t ez = cudaMemset(h->k.reset_list, 0, sizeof(int) * n_int);
  if (ez == cudaSuccess) ez = cudaMemset(h->k.stats, 0, sizeof(double) * 5);
  if (ez == cudaSuccess) { ops.init(h->k, 0); ez = cudaGetLastError(); }
  if (ez == cudaSuccess) ez = cudaStreamSynchronize(0);
  if (ez != cudaSuccess) CREATE_FAIL(XARM_E_CUDA, std::string("xarm_create: ") + cudaGetErrorString(ez));
#undef CREATE_FAIL
  *out = h;
  return XARM_OK;
}

int xarm_destroy(XarmHandle* h) {
  if (!h) return XARM_OK;
  cudaSetDevice(h->cfg.device);
  if (h->graph) cudaGraphExecDestroy(h->graph);
  if (h->graph_host) cudaGraphExecDestroy(h->graph_host);
  if (h->host_stream) cudaStreamDestroy(h->host_stream);
  cudaFree(h->d_term);
  if (h->h_term) cudaFreeHost(h->h_term);
  cudaFree(h->k.state); cudaFree(h->k.ep_return); cudaFree(h->k.need_reset); cudaFree(h->k.stats); cudaFree(h->k.reset_list); cudaFree(h->k.scratch);
  cudaFree(h->pipe.hrec); cudaFree(h->pipe.tl_dev);
  for (cudaEvent_t e : h->pipe.ev) cudaEventDestroy(e);
  if (h->pipe.side) cudaStreamDestroy(h->pipe.side);
  if (h->pipe.e_main) cudaStreamDestroy(h->pipe.e_main);
  if (h->pipe.e_side) cudaStreamDestroy(h->pipe.e_side);
  if (h->pipe.e_side2) cudaStreamDestroy(h->pipe.e_side2);
  cudaFree(h->d_io); cudaFree(h->d_flags);
  if (h->h_io) cudaFreeHost(h->h_io);
  if (h->h_flags) cudaFreeHost(h->h_flags);
  delete h;
  return XARM_OK;
}

int xarm_bind(XarmHandle* h, const XarmBuffers* b) {
  if (!h || !b) return fail(XARM_E_INVALID, "xarm_bind: null argument");
  if (!b->actions || !b->observation || !b->achieved_goal || !b->desired_goal || !b->reward || !b->done || !b->success)
    return fail(XARM_E_INVALID, "xarm_bind: actions/observation/achieved_goal/desired_goal/reward/done/success are required");
  h->k.b = *b; h->user_bufs = *b; h->user_bound = true;
  h->bound = true;
  if (h->graph) { cudaGraphExecDestroy(h->graph); h->graph = nullptr; }  // addresses changed: recapture
  return XARM_OK;
}

// capture launch_step_with(h, k, s) into *exec (shared by the device-buffer and the host-buffer paths)
static int capture_step(XarmHandle* h, const KArgs& k, cudaStream_t s, bool with_gather, cudaGraphExec_t* exec, int64_t* n_launches);

// one env step = the kernel pipeline of xarm_pipeline.cuh (action -> NSUB x {setup -> light || heavy} -> finish -> auto-reset tail)
static int launch_step_with(XarmHandle* h, const KArgs& k, cudaStream_t s) {
  h->ops.step(h->pipe, k, s);
  return XARM_OK;
}
static int launch_step(XarmHandle* h, cudaStream_t s) { return launch_step_with(h, h->k, s); }

int xarm_reset(XarmHandle* h, const uint8_t* mask, void* stream) {
  if (!h) return fail(XARM_E_INVALID, "xarm_reset: null handle");
  if (!h->bound) return fail(XARM_E_STATE, "xarm_reset: call xarm_bind first");
  CUDA_TRY(cudaSetDevice(h->cfg.device));
  h->ops.reset(h->pipe, h->k, mask, (cudaStream_t)stream);
  CUDA_TRY(cudaGetLastError());
  return XARM_OK;
}

int xarm_step(XarmHandle* h, void* stream) {
  if (!h) return fail(XARM_E_INVALID, "xarm_step: null handle");
  if (!h->bound) return fail(XARM_E_STATE, "xarm_step: call xarm_bind first");
  CUDA_TRY(cudaSetDevice(h->cfg.device));
  cudaStream_t s = (cudaStream_t)stream;
  if (h->graph && s == h->graph_stream) {
    CUDA_TRY(cudaGraphLaunch(h->graph, s));
    g_launches += h->graph_launches;
  } else {
    launch_step(h, s);
    CUDA_TRY(cudaGetLastError());
    h->pipe.report();
  }
  return XARM_OK;
}

static void launch_gather_terminal(XarmHandle* h, cudaStream_t s);
static int capture_step(XarmHandle* h, const KArgs& k, cudaStream_t s, bool with_gather, cudaGraphExec_t* exec, int64_t* n_launches) {
  if (*exec) { cudaGraphExecDestroy(*exec); *exec = nullptr; }
  cudaGraph_t g = nullptr;
  int64_t before = g_launches.load();
  CUDA_TRY(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
  launch_step_with(h, k, s);
  if (with_gather) launch_gather_terminal(h, s);
  cudaError_t e = cudaStreamEndCapture(s, &g);
  *n_launches = g_launches.load() - before;
  g_launches = before;  // captured launches did not run
  if (e != cudaSuccess) return fail(XARM_E_CUDA, std::string("cudaStreamEndCapture: ") + cudaGetErrorString(e));
  e = cudaGraphInstantiate(exec, g, 0);
  cudaGraphDestroy(g);
  if (e != cudaSuccess) { *exec = nullptr; return fail(XARM_E_CUDA, std::string("cudaGraphInstantiate: ") + cudaGetErrorString(e)); }
  return XARM_OK;
}

int xarm_graph_capture(XarmHandle* h, void* stream) {
  if (!h) return fail(XARM_E_INVALID, "xarm_graph_capture: null handle");
  if (!h->bound) return fail(XARM_E_STATE, "xarm_graph_capture: call xarm_bind first");
  CUDA_TRY(cudaSetDevice(h->cfg.device));
  cudaStream_t s = (cudaStream_t)stream;
  if (s == nullptr) return fail(XARM_E_INVALID, "xarm_graph_capture: needs a non-default stream");
  int rc = capture_step(h, h->k, s, false, &h->graph, &h->graph_launches);
  if (rc) return rc;
  h->graph_stream = s;
  return XARM_OK;
}

static int ensure_host_io(XarmHandle* h) {
  if (h->d_io) return XARM_OK;
  const int64_t n = h->cfg.num_envs;
  const Ops& o = h->ops;
  const int W = o.O + 2 * o.G;
  const size_t fl = (size_t)n * (o.A + W + 2), fl_dev = fl + (size_t)2 * n * W;   // + terminal slab + gathered rows
  CUDA_TRY(cudaMalloc(&h->d_io, fl_dev * sizeof(float)));
  CUDA_TRY(cudaMalloc(&h->d_flags, 2 * n));
  CUDA_TRY(cudaMalloc(&h->d_term, sizeof(int) * (1 + n)));
  CUDA_TRY(cudaMallocHost(&h->h_io, (fl + (size_t)n * W) * sizeof(float)));
  CUDA_TRY(cudaMallocHost(&h->h_flags, 2 * n));
  CUDA_TRY(cudaMallocHost(&h->h_term, sizeof(int) * (1 + n)));
  CUDA_TRY(cudaStreamCreateWithFlags(&h->host_stream, cudaStreamNonBlocking));
  XarmBuffers& b = h->host_bufs;
  float* p = h->d_io;
  b.actions = p; p += n * o.A;
  b.observation = p; p += n * o.O;
  b.achieved_goal = p; p += n * o.G;
  b.desired_goal = p; p += n * o.G;
  b.reward = p; p += n;
  b.success = p; p += n;
  b.terminal_observation = p;   // [N, W]; the gathered rows follow at p + n * W
  b.done = h->d_flags; b.truncated = h->d_flags + n;
  return XARM_OK;
}
static void launch_gather_terminal(XarmHandle* h, cudaStream_t s) {
  const int64_t n = h->cfg.num_envs;
  const int W = h->ops.O + 2 * h->ops.G;
  cudaMemsetAsync(h->d_term, 0, sizeof(int), s);
  const unsigned grid = (unsigned)((n + 255) / 256 < 592 ? (n + 255) / 256 : 592);
  k_gather_terminal<<<grid, 256, 0, s>>>(h->host_bufs.done, h->host_bufs.terminal_observation, n, W, h->host_bufs.terminal_observation + n * W, h->d_term + 1, h->d_term);
  g_launches++;
}

// The numpy-facing path: host buffers in, host buffers out.  Uses the library's own device staging; pinned host
// staging keeps the copies asynchronous with respect to the step kernels.  The step itself is a replay of a CUDA graph
// captured at the first call (XARM_HOST_GRAPH=0: plain launches).
int xarm_step_host(XarmHandle* h, const float* actions, float* observation, float* achieved_goal, float* desired_goal,
                   float* reward, uint8_t* done, float* success, uint8_t* truncated, float* terminal_observation, void* stream) {
  if (!h || !actions) return fail(XARM_E_INVALID, "xarm_step_host: null argument");
  CUDA_TRY(cudaSetDevice(h->cfg.device));
  int rc = ensure_host_io(h);
  if (rc) return rc;
  cudaStream_t s = stream ? (cudaStream_t)stream : h->host_stream;
  const int64_t n = h->cfg.num_envs;
  const Ops& o = h->ops;
  const int W = o.O + 2 * o.G;
  KArgs k = h->k;
  k.b = h->host_bufs;
  float* hp = h->h_io;
  memcpy(hp, actions, sizeof(float) * n * o.A);
  CUDA_TRY(cudaMemcpyAsync((void*)k.b.actions, hp, sizeof(float) * n * o.A, cudaMemcpyHostToDevice, s));
  static const bool use_graph = !(getenv("XARM_HOST_GRAPH") && atoi(getenv("XARM_HOST_GRAPH")) == 0);
  if (use_graph && !h->pipe.trace) {
    if (!h->graph_host || h->graph_host_stream != s) {
      CUDA_TRY(cudaStreamSynchronize(s));
      rc = capture_step(h, k, s, true, &h->graph_host, &h->graph_host_launches);
      if (rc) return rc;
      h->graph_host_stream = s;
    }
    CUDA_TRY(cudaGraphLaunch(h->graph_host, s));
    g_launches += h->graph_host_launches;
  } else {
    launch_step_with(h, k, s);
    launch_gather_terminal(h, s);
    CUDA_TRY(cudaGetLastError());
  }
  // one D2H copy of the contiguous float outputs, one of the flags, one of the finished-env count
  const size_t out_fl = (size_t)n * (W + 2);
  CUDA_TRY(cudaMemcpyAsync(hp + n * o.A, k.b.observation, out_fl * sizeof(float), cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaMemcpyAsync(h->h_flags, h->d_flags, 2 * n, cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaMemcpyAsync(h->h_term, h->d_term, sizeof(int), cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaStreamSynchronize(s));
  const float* q = hp + n * o.A;
  if (observation) memcpy(observation, q, sizeof(float) * n * o.O);
  q += n * o.O;
  if (achieved_goal) memcpy(achieved_goal, q, sizeof(float) * n * o.G);
  q += n * o.G;
  if (desired_goal) memcpy(desired_goal, q, sizeof(float) * n * o.G);
  q += n * o.G;
  if (reward) memcpy(reward, q, sizeof(float) * n);
  q += n;
  if (success) memcpy(success, q, sizeof(float) * n);
  if (done) memcpy(done, h->h_flags, n);
  if (truncated) memcpy(truncated, h->h_flags + n, n);
  const int n_term = h->h_term[0];
  if (terminal_observation && n_term > 0) {   // second, small transfer: the rows of the envs that finished + their ids
    float* hrows = hp + (size_t)n * (o.A + W + 2);
    CUDA_TRY(cudaMemcpyAsync(hrows, k.b.terminal_observation + n * W, sizeof(float) * (size_t)n_term * W, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaMemcpyAsync(h->h_term + 1, h->d_term + 1, sizeof(int) * n_term, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    for (int r = 0; r < n_term; r++) memcpy(terminal_observation + (size_t)h->h_term[1 + r] * W, hrows + (size_t)r * W, sizeof(float) * W);
  }
  return XARM_OK;
}

int xarm_reset_host(XarmHandle* h, float* observation, float* achieved_goal, float* desired_goal, void* stream) {
  if (!h) return fail(XARM_E_INVALID, "xarm_reset_host: null handle");
  CUDA_TRY(cudaSetDevice(h->cfg.device));
  int rc = ensure_host_io(h);
  if (rc) return rc;
  cudaStream_t s = stream ? (cudaStream_t)stream : h->host_stream;
  const int64_t n = h->cfg.num_envs;
  const Ops& o = h->ops;
  KArgs k = h->k;
  k.b = h->host_bufs;
  h->ops.reset(h->pipe, k, nullptr, s);
  CUDA_TRY(cudaGetLastError());
  float* hp = h->h_io + n * o.A;
  CUDA_TRY(cudaMemcpyAsync(hp, k.b.observation, sizeof(float) * n * (o.O + 2 * o.G), cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaStreamSynchronize(s));
  if (observation) memcpy(observation, hp, sizeof(float) * n * o.O);
  if (achieved_goal) memcpy(achieved_goal, hp + n * o.O, sizeof(float) * n * o.G);
  if (desired_goal) memcpy(desired_goal, hp + n * (o.O + o.G), sizeof(float) * n * o.G);
  return XARM_OK;
}

int xarm_compute_reward(int32_t task, int32_t reward_type, int32_t num_obj, const float* ag, const float* dg, int64_t n,
                        float* out, void* stream) {
  Ops o;
  num_obj = norm_num_obj(task, num_obj);
  if (task < 0 || task >= XARM_NUM_TASKS) return fail(XARM_E_INVALID, "xarm_compute_reward: bad task");
  int G = task == XARM_TASK_REACH ? 3 : 3 * num_obj;
  if (G < 3 || G > 9) return fail(XARM_E_INVALID, "xarm_compute_reward: num_obj out of range");
  bool ok = reward_type == XARM_REWARD_SPARSE ||
            (reward_type == XARM_REWARD_DENSE && (task == XARM_TASK_REACH || task == XARM_TASK_STACK_TOWER || task == XARM_TASK_PUSH_WITH_DOOR)) ||
            (reward_type == XARM_REWARD_DENSE_O2G && task == XARM_TASK_PICK_AND_PLACE);
  if (!ok) return fail(XARM_E_INVALID, "xarm_compute_reward: this reward type reads simulator state (not batch-safe in the reference either)");
  if (n <= 0) return XARM_OK;
  if (!ag || !dg || !out) return fail(XARM_E_INVALID, "xarm_compute_reward: null pointer");
  (void)o;
  k_compute_reward<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(task, reward_type, num_obj, G, ag, dg, n, out);
  g_launches++;
  CUDA_TRY(cudaGetLastError());
  return XARM_OK;
}

int xarm_get_state(XarmHandle* h, float* host_out) {
  if (!h || !host_out) return fail(XARM_E_INVALID, "xarm_get_state: null argument");
  CUDA_TRY(cudaSetDevice(h->cfg.device));
  const int64_t n = h->cfg.num_envs; const int S = h->ops.S;
  std::vector<float> tmp((size_t)n * S);
  CUDA_TRY(cudaDeviceSynchronize());
  CUDA_TRY(cudaMemcpy(tmp.data(), h->k.state, sizeof(float) * n * S, cudaMemcpyDeviceToHost));
  for (int64_t i = 0; i < n; i++)
    for (int w = 0; w < S; w++) host_out[i * S + w] = tmp[(size_t)w * n + i];
  return XARM_OK;
}

int xarm_set_state(XarmHandle* h, const float* host_in) {
  if (!h || !host_in) return fail(XARM_E_INVALID, "xarm_set_state: null argument");
  CUDA_TRY(cudaSetDevice(h->cfg.device));
  const int64_t n = h->cfg.num_envs; const int S = h->ops.S;
  std::vector<float> tmp((size_t)n * S);
  for (int64_t i = 0; i < n; i++)
    for (int w = 0; w < S; w++) tmp[(size_t)w * n + i] = host_in[i * S + w];
  CUDA_TRY(cudaDeviceSynchronize());
  CUDA_TRY(cudaMemcpy(h->k.state, tmp.data(), sizeof(float) * n * S, cudaMemcpyHostToDevice));
  return XARM_OK;
}

int xarm_get_obs(XarmHandle* h, void* stream) {
  if (!h) return fail(XARM_E_INVALID, "xarm_get_obs: null handle");
  if (!h->bound) return fail(XARM_E_STATE, "xarm_get_obs: call xarm_bind first");
  CUDA_TRY(cudaSetDevice(h->cfg.device));
  h->ops.obs(h->k, (cudaStream_t)stream);
  CUDA_TRY(cudaGetLastError());
  return XARM_OK;
}

// Profiling aid of bench.py: %globaltimer stamps around every pipeline launch (first block start .. last block end),
// accumulated on the device per launch slot - it works inside the captured graph and in the real concurrent schedule.
int xarm_set_profiling(XarmHandle* h, int32_t on) {
  if (!h) return fail(XARM_E_INVALID, "xarm_set_profiling: null handle");
  CUDA_TRY(cudaSetDevice(h->cfg.device));
  if (on && !h->pipe.tl_dev) CUDA_TRY(cudaMalloc(&h->pipe.tl_dev, sizeof(unsigned long long) * (5 * XARM_TL_SLOTS + 8)));
  if (h->pipe.tl_dev) CUDA_TRY(cudaMemset(h->pipe.tl_dev, 0, sizeof(unsigned long long) * (5 * XARM_TL_SLOTS + 8)));
  CUDA_TRY(cudaDeviceSynchronize());
  const bool was = h->pipe.timeline;
  h->pipe.timeline = on != 0;   // (calling it again while on just clears the accumulators)
  if (was != h->pipe.timeline) {   // the launches carry the buffer: recapture
    if (h->graph) { cudaGraphExecDestroy(h->graph); h->graph = nullptr; }
    if (h->graph_host) { cudaGraphExecDestroy(h->graph_host); h->graph_host = nullptr; }
  }
  return XARM_OK;
}

// text lines "branch kernel launches total_us" summed over the steps since xarm_set_profiling (branch: M main, E early
// = envs that may finish + their auto-reset passes, L late tail); returns the number of lines, < 0 on error
int xarm_kernel_times(XarmHandle* h, char* out, int64_t cap) {
  if (!h || !out || cap < 1) return fail(XARM_E_INVALID, "xarm_kernel_times: null argument");
  if (!h->pipe.timeline || !h->pipe.tl_dev) return fail(XARM_E_STATE, "xarm_kernel_times: call xarm_set_profiling(h, 1) first");
  CUDA_TRY(cudaSetDevice(h->cfg.device));
  CUDA_TRY(cudaDeviceSynchronize());
  const size_t nslot = h->pipe.tl_names.size();
  std::vector<unsigned long long> t(5 * XARM_TL_SLOTS + 8);
  CUDA_TRY(cudaMemcpy(t.data(), h->pipe.tl_dev, sizeof(unsigned long long) * (5 * XARM_TL_SLOTS + 8), cudaMemcpyDeviceToHost));
  std::vector<std::pair<std::string, std::pair<unsigned long long, unsigned long long>>> acc;
  for (size_t k = 0; k < nslot; k++) {
    unsigned long long ns = t[3 * XARM_TL_SLOTS + k], cnt = t[4 * XARM_TL_SLOTS + k];
    if (t[2 * k] != ~0ull && t[2 * k + 1] >= t[2 * k]) { ns += t[2 * k + 1] - t[2 * k]; cnt += 1; }  // the last step
    size_t j = 0;
    for (; j < acc.size(); j++) if (acc[j].first == h->pipe.tl_names[k]) break;
    if (j == acc.size()) acc.push_back({h->pipe.tl_names[k], {0ull, 0ull}});
    acc[j].second.first += cnt; acc[j].second.second += ns;
  }
  std::string txt;
  char line[160];
  for (auto& kv : acc) {
    snprintf(line, sizeof(line), "%s %llu %.1f\n", kv.first.c_str(), kv.second.first, kv.second.second * 1e-3);
    txt += line;
  }
  // env-substep counters per branch: "#heavy_envs" = envs that took the cooperative contact path, "#setup_envs" = all
  for (int b = 0; b < 3; b++) {
    snprintf(line, sizeof(line), "%c #heavy_envs %llu 0\n%c #setup_envs %llu 0\n", "MEL"[b], t[5 * XARM_TL_SLOTS + b], "MEL"[b], t[5 * XARM_TL_SLOTS + 3 + b]);
    txt += line;
  }
  if ((int64_t)txt.size() + 1 > cap) txt.resize(cap - 1);
  memcpy(out, txt.c_str(), txt.size() + 1);
  return (int)acc.size();
}

// development aid (not part of include/xarm_abi.h): timeline of the last step as text lines "branch kernel start_us end_us"
int xarm_debug_timeline(XarmHandle* h, char* out, int64_t cap) {
  if (!h || !out || !h->pipe.timeline) return -1;
  cudaSetDevice(h->cfg.device);
  cudaDeviceSynchronize();
  const size_t nslot = h->pipe.tl_names.size();
  std::vector<unsigned long long> t(3 * XARM_TL_SLOTS);
  cudaMemcpy(t.data(), h->pipe.tl_dev, sizeof(unsigned long long) * 3 * XARM_TL_SLOTS, cudaMemcpyDeviceToHost);
  unsigned long long t0 = ~0ull;
  for (size_t k = 0; k < nslot; k++) if (t[2 * k] < t0) t0 = t[2 * k];
  std::string txt;
  char line[160];
  for (size_t k = 0; k < nslot; k++) {
    if (t[2 * k] == ~0ull) continue;  // launch had no work
    snprintf(line, sizeof(line), "%s %.1f %.1f %llu\n", h->pipe.tl_names[k].c_str(), (t[2 * k] - t0) * 1e-3, (t[2 * k + 1] - t0) * 1e-3, t[2 * XARM_TL_SLOTS + k]);
    txt += line;
  }
  if ((int64_t)txt.size() + 1 > cap) txt.resize(cap - 1);
  memcpy(out, txt.c_str(), txt.size() + 1);
  return (int)nslot;
}

// development aid (not part of include/xarm_abi.h): the heavy-solver records of the last substep, [N][HeavyRec::WORDS]
int xarm_debug_hrec(XarmHandle* h, float* host_out, int64_t max_words) {
  if (!h || !host_out || !h->pipe.hrec) return -1;
  cudaSetDevice(h->cfg.device);
  cudaDeviceSynchronize();
  const int64_t words = (int64_t)h->ops.hrec_words() * h->cfg.num_envs;
  cudaMemcpy(host_out, h->pipe.hrec, sizeof(float) * (words < max_words ? words : max_words), cudaMemcpyDeviceToHost);
  return (int)h->ops.hrec_words();
}

int xarm_episode_stats(XarmHandle* h, double out[5], void* stream) {
  if (!h || !out) return fail(XARM_E_INVALID, "xarm_episode_stats: null argument");
  CUDA_TRY(cudaSetDevice(h->cfg.device));
  cudaStream_t s = (cudaStream_t)stream;
  CUDA_TRY(cudaMemcpyAsync(out, h->k.stats, sizeof(double) * 5, cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaMemsetAsync(h->k.stats, 0, sizeof(double) * 5, s));
  CUDA_TRY(cudaStreamSynchronize(s));
  return XARM_OK;
}

// FP32 SIMT peak of the device, measured: 8 independent FMA chains per thread, full occupancy
__global__ void __launch_bounds__(256) k_fma_chain(float* out, int iters, float a, float b) {
  float x0 = threadIdx.x, x1 = x0 + 1.f, x2 = x0 + 2.f, x3 = x0 + 3.f, x4 = x0 + 4.f, x5 = x0 + 5.f, x6 = x0 + 6.f, x7 = x0 + 7.f;
#pragma unroll 1
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int u = 0; u < 8; u++) {
      x0 = fmaf(x0, a, b); x1 = fmaf(x1, a, b); x2 = fmaf(x2, a, b); x3 = fmaf(x3, a, b);
      x4 = fmaf(x4, a, b); x5 = fmaf(x5, a, b); x6 = fmaf(x6, a, b); x7 = fmaf(x7, a, b);
    }
  }
  const float s = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
  if (s == 123.456f) out[blockIdx.x * blockDim.x + threadIdx.x] = s;   // never true: keeps the chains alive
}
int xarm_measure_fp32_peak(int32_t device, double* tflops) {
  if (!tflops) return fail(XARM_E_INVALID, "xarm_measure_fp32_peak: null argument");
  int ndev = 0;
  CUDA_TRY(cudaGetDeviceCount(&ndev));
  if (device < 0 || device >= ndev) return fail(XARM_E_INVALID, "xarm_measure_fp32_peak: bad device ordinal");
  CUDA_TRY(cudaSetDevice(device));
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  float* out = nullptr;
  const unsigned grid = (unsigned)sms * 8;   // 8 x 256 threads per SM: all 64 warp slots
  CUDA_TRY(cudaMalloc(&out, sizeof(float) * grid * 256));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 4096;   // x 64 FMAs per thread and iteration
  double best = 0.0;
  for (int rep = 0; rep < 6; rep++) {
    cudaEventRecord(e0, 0);
    k_fma_chain<<<grid, 256>>>(out, iters, 0.999f, 0.001f);
    cudaEventRecord(e1, 0);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    const double tf = 2.0 * 64.0 * iters * (double)grid * 256.0 / (ms * 1e-3) / 1e12;
    if (rep > 0 && tf > best) best = tf;
  }
  g_launches += 6;
  cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(out);
  CUDA_TRY(cudaGetLastError());
  *tflops = best;
  return XARM_OK;
}

// ------------------------------------------------------------------------------------------------ VecNormalize
struct XarmVecNorm {
  XarmVecNormConfig cfg;
  VnStats* st = nullptr;     // device
  float* ret = nullptr;      // [N] discounted returns
  unsigned grid = 0;         // grid * 256 threads: a multiple of obs_dim
};

static int vn_gcd(int a, int b) { while (b) { int t = a % b; a = b; b = t; } return a; }

int xarm_vecnorm_create(const XarmVecNormConfig* cfg, XarmVecNorm** out) {
  if (!cfg || !out) return fail(XARM_E_INVALID, "xarm_vecnorm_create: null argument");
  if (cfg->num_envs <= 0 || cfg->obs_dim <= 0 || cfg->obs_dim > XARM_VN_MAX_OBS) return fail(XARM_E_INVALID, "xarm_vecnorm_create: num_envs > 0 and 0 < obs_dim <= 128 required");
  int ndev = 0;
  CUDA_TRY(cudaGetDeviceCount(&ndev));
  if (cfg->device < 0 || cfg->device >= ndev) return fail(XARM_E_INVALID, "xarm_vecnorm_create: bad device ordinal");
  CUDA_TRY(cudaSetDevice(cfg->device));
  XarmVecNorm* v = new (std::nothrow) XarmVecNorm();
  if (!v) return fail(XARM_E_NOMEM, "xarm_vecnorm_create: out of host memory");
  v->cfg = *cfg;
  if (cudaMalloc(&v->st, sizeof(VnStats)) != cudaSuccess || cudaMalloc(&v->ret, sizeof(float) * cfg->num_envs) != cudaSuccess) {
    cudaFree(v->st); cudaFree(v->ret); delete v; cudaGetLastError();
    return fail(XARM_E_NOMEM, "xarm_vecnorm_create: cudaMalloc failed");
  }
  VnStats h;
  memset(&h, 0, sizeof(h));
  for (int k = 0; k < XARM_VN_MAX_OBS; k++) h.var[k] = 1.0;     // RunningMeanStd(): mean 0, var 1, count 1e-4
  h.count = 1e-4; h.rvar = 1.0; h.rcount = 1e-4;
  CUDA_TRY(cudaMemcpy(v->st, &h, sizeof(h), cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMemset(v->ret, 0, sizeof(float) * cfg->num_envs));
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, cfg->device);
  const int unit = cfg->obs_dim / vn_gcd(cfg->obs_dim, 256);     // grid must be a multiple of this
  int64_t want = (cfg->num_envs * cfg->obs_dim + 256 * 8 - 1) / (256 * 8);   // ~8 elements per thread
  if (want > (int64_t)sms * 8) want = (int64_t)sms * 8;
  if (want < 1) want = 1;
  v->grid = (unsigned)((want + unit - 1) / unit * unit);
  k_vn_finalize<<<1, XARM_VN_MAX_OBS>>>(v->st, 1, cfg->obs_dim, cfg->epsilon, 0, 0);  // fmean / finv / rinv of the initial statistics
  g_launches++;
  CUDA_TRY(cudaDeviceSynchronize());
  *out = v;
  return XARM_OK;
}

int xarm_vecnorm_destroy(XarmVecNorm* v) {
  if (!v) return XARM_OK;
  cudaSetDevice(v->cfg.device);
  cudaFree(v->st); cudaFree(v->ret);
  delete v;
  return XARM_OK;
}

int xarm_vecnorm_set_training(XarmVecNorm* v, int32_t training) {
  if (!v) return fail(XARM_E_INVALID, "xarm_vecnorm_set_training: null handle");
  v->cfg.training = training;
  return XARM_OK;
}

int xarm_vecnorm_reset(XarmVecNorm* v, const float* obs, float* obs_out, void* stream) {
  if (!v || !obs || !obs_out) return fail(XARM_E_INVALID, "xarm_vecnorm_reset: null argument");
  CUDA_TRY(cudaSetDevice(v->cfg.device));
  cudaStream_t s = (cudaStream_t)stream;
  const XarmVecNormConfig& c = v->cfg;
  // returns := 0; ret_rms.update(returns) when training (stable-baselines3 1.x); obs_rms is not touched
  k_vn_moments<<<v->grid, 256, 0, s>>>(v->st, obs, c.num_envs, c.obs_dim, v->ret, nullptr, c.gamma, 0, 2);
  k_vn_finalize<<<1, XARM_VN_MAX_OBS, 0, s>>>(v->st, c.num_envs, c.obs_dim, c.epsilon, 0, c.training != 0);
  k_vn_apply<<<v->grid, 256, 0, s>>>(v->st, obs, obs_out, c.num_envs, c.obs_dim, nullptr, nullptr, nullptr, v->ret, c.clip_obs, c.clip_reward, c.norm_obs, c.norm_reward);
  g_launches += 3;
  CUDA_TRY(cudaGetLastError());
  return XARM_OK;
}

int xarm_vecnorm_step(XarmVecNorm* v, const float* obs, const float* reward, const uint8_t* done, float* obs_out, float* reward_out, void* stream) {
  if (!v || !obs || !reward || !done || !obs_out || !reward_out) return fail(XARM_E_INVALID, "xarm_vecnorm_step: null argument");
  CUDA_TRY(cudaSetDevice(v->cfg.device));
  cudaStream_t s = (cudaStream_t)stream;
  const XarmVecNormConfig& c = v->cfg;
  const int tr = c.training != 0;
  if (tr) { k_vn_moments<<<v->grid, 256, 0, s>>>(v->st, obs, c.num_envs, c.obs_dim, v->ret, reward, c.gamma, c.norm_obs != 0, 1); g_launches++; }
  k_vn_finalize<<<1, XARM_VN_MAX_OBS, 0, s>>>(v->st, c.num_envs, c.obs_dim, c.epsilon, tr && c.norm_obs, tr);
  k_vn_apply<<<v->grid, 256, 0, s>>>(v->st, obs, obs_out, c.num_envs, c.obs_dim, reward, reward_out, done, v->ret, c.clip_obs, c.clip_reward, c.norm_obs, c.norm_reward);
  g_launches += 2;
  CUDA_TRY(cudaGetLastError());
  return XARM_OK;
}

int xarm_vecnorm_normalize_obs(XarmVecNorm* v, const float* obs, int64_t n, int64_t row_stride, float* out, int64_t out_stride, int32_t inverse, void* stream) {
  if (!v || !obs || !out) return fail(XARM_E_INVALID, "xarm_vecnorm_normalize_obs: null argument");
  const XarmVecNormConfig& c = v->cfg;
  if (n < 0 || row_stride < c.obs_dim || out_stride < c.obs_dim) return fail(XARM_E_INVALID, "xarm_vecnorm_normalize_obs: n >= 0 and strides >= obs_dim required");
  if (n == 0) return XARM_OK;
  CUDA_TRY(cudaSetDevice(c.device));
  const int64_t blocks = (n * c.obs_dim + 255) / 256;
  k_vn_rows<<<(unsigned)(blocks < 4736 ? blocks : 4736), 256, 0, (cudaStream_t)stream>>>(v->st, obs, n, c.obs_dim, row_stride, out, out_stride, c.clip_obs, c.epsilon, c.norm_obs, inverse);
  g_launches++;
  CUDA_TRY(cudaGetLastError());
  return XARM_OK;
}

int xarm_vecnorm_normalize_reward(XarmVecNorm* v, const float* reward, int64_t n, float* out, void* stream) {
  if (!v || !reward || !out) return fail(XARM_E_INVALID, "xarm_vecnorm_normalize_reward: null argument");
  if (n < 0) return fail(XARM_E_INVALID, "xarm_vecnorm_normalize_reward: n >= 0 required");
  if (n == 0) return XARM_OK;
  const XarmVecNormConfig& c = v->cfg;
  CUDA_TRY(cudaSetDevice(c.device));
  const int64_t blocks = (n + 255) / 256;
  k_vn_reward_rows<<<(unsigned)(blocks < 4736 ? blocks : 4736), 256, 0, (cudaStream_t)stream>>>(v->st, reward, n, out, c.clip_reward, c.norm_reward);
  g_launches++;
  CUDA_TRY(cudaGetLastError());
  return XARM_OK;
}

int xarm_vecnorm_get_stats(XarmVecNorm* v, double* obs_mean, double* obs_var, double* obs_count, double* ret_stats3) {
  if (!v) return fail(XARM_E_INVALID, "xarm_vecnorm_get_stats: null handle");
  CUDA_TRY(cudaSetDevice(v->cfg.device));
  CUDA_TRY(cudaDeviceSynchronize());
  VnStats h;
  CUDA_TRY(cudaMemcpy(&h, v->st, sizeof(h), cudaMemcpyDeviceToHost));
  for (int k = 0; k < v->cfg.obs_dim; k++) { if (obs_mean) obs_mean[k] = h.mean[k]; if (obs_var) obs_var[k] = h.var[k]; }
  if (obs_count) *obs_count = h.count;
  if (ret_stats3) { ret_stats3[0] = h.rmean; ret_stats3[1] = h.rvar; ret_stats3[2] = h.rcount; }
  return XARM_OK;
}

int xarm_vecnorm_set_stats(XarmVecNorm* v, const double* obs_mean, const double* obs_var, double obs_count, const double* ret_stats3) {
  if (!v || !obs_mean || !obs_var || !ret_stats3) return fail(XARM_E_INVALID, "xarm_vecnorm_set_stats: null argument");
  CUDA_TRY(cudaSetDevice(v->cfg.device));
  CUDA_TRY(cudaDeviceSynchronize());
  VnStats h;
  CUDA_TRY(cudaMemcpy(&h, v->st, sizeof(h), cudaMemcpyDeviceToHost));
  for (int k = 0; k < v->cfg.obs_dim; k++) { h.mean[k] = obs_mean[k]; h.var[k] = obs_var[k]; }
  h.count = obs_count; h.rmean = ret_stats3[0]; h.rvar = ret_stats3[1]; h.rcount = ret_stats3[2];
  CUDA_TRY(cudaMemcpy(v->st, &h, sizeof(h), cudaMemcpyHostToDevice));
  k_vn_finalize<<<1, XARM_VN_MAX_OBS>>>(v->st, 1, v->cfg.obs_dim, v->cfg.epsilon, 0, 0);
  g_launches++;
  CUDA_TRY(cudaDeviceSynchronize());
  return XARM_OK;
}

// ------------------------------------------------------------------------------------------------ hindsight experience replay
struct XarmHer {
  XarmHerConfig cfg;
  HerBuf b;
  int4* index = nullptr;        // scratch of the last sample call
  int64_t index_cap = 0;
  uint32_t calls = 0;
};

static void her_free(XarmHer* h) {
  cudaFree(h->b.rec); cudaFree(h->b.dg); cudaFree(h->b.ep_len); cudaFree(h->b.cur_k); cudaFree(h->b.cur_t); cudaFree(h->b.counters);
  cudaFree(h->index);
  cudaGetLastError();
}

int xarm_her_create(const XarmHerConfig* cfg, XarmHer** out) {
  if (!cfg || !out) return fail(XARM_E_INVALID, "xarm_her_create: null argument");
  if (cfg->num_envs <= 0 || cfg->num_envs > 0x7fffffffLL) return fail(XARM_E_INVALID, "xarm_her_create: 0 < num_envs < 2^31 required");
  if (cfg->episodes_per_env < 2) return fail(XARM_E_INVALID, "xarm_her_create: episodes_per_env >= 2 required (one slot is always being written)");
  if (cfg->max_episode_length < 1) return fail(XARM_E_INVALID, "xarm_her_create: max_episode_length >= 1 required");
  if (cfg->obs_dim < 1 || cfg->obs_dim > XARM_HER_MAX_OBS || cfg->action_dim < 1 || cfg->action_dim > XARM_HER_MAX_ACT || cfg->goal_dim < 1 || cfg->goal_dim > 9)
    return fail(XARM_E_INVALID, "xarm_her_create: 1 <= obs_dim <= 128, 1 <= action_dim <= 32 and 1 <= goal_dim <= 9 required");
  if (cfg->n_sampled_goal < 0) return fail(XARM_E_INVALID, "xarm_her_create: n_sampled_goal >= 0 required");
  if ((double)cfg->episodes_per_env * ((double)cfg->max_episode_length + 1.0) * (double)cfg->num_envs >= 4294967296.0)
    return fail(XARM_E_INVALID, "xarm_her_create: episodes_per_env x (max_episode_length + 1) x num_envs must stay below 2^32 rows");
  if (cfg->task < 0 || cfg->task >= XARM_NUM_TASKS) return fail(XARM_E_INVALID, "xarm_her_create: bad task");
  {
    const int rt = cfg->reward_type;   // the same gate as xarm_compute_reward: relabelling needs a state-free reward
    const bool ok = rt == XARM_REWARD_SPARSE || (rt == XARM_REWARD_DENSE && cfg->task != XARM_TASK_PICK_AND_PLACE && cfg->task != XARM_TASK_HANDOVER) ||
                    (rt == XARM_REWARD_DENSE_O2G && cfg->task == XARM_TASK_PICK_AND_PLACE);
    if (!ok) return fail(XARM_E_INVALID, "xarm_her_create: this reward type reads simulator state (not batch-safe in the reference either)");
  }
  int ndev = 0;
  CUDA_TRY(cudaGetDeviceCount(&ndev));
  if (cfg->device < 0 || cfg->device >= ndev) return fail(XARM_E_INVALID, "xarm_her_create: bad device ordinal");
  CUDA_TRY(cudaSetDevice(cfg->device));
  XarmHer* h = new (std::nothrow) XarmHer();
  if (!h) return fail(XARM_E_NOMEM, "xarm_her_create: out of host memory");
  h->cfg = *cfg;
  HerBuf& b = h->b;
  memset(&b, 0, sizeof(b));
  b.N = cfg->num_envs; b.K = cfg->episodes_per_env; b.T = cfg->max_episode_length;
  b.O = cfg->obs_dim; b.G = cfg->goal_dim; b.A = cfg->action_dim;
  const size_t recs = (size_t)b.K * (b.T + 1) * b.N, eps = (size_t)b.K * b.N, rw = (size_t)(b.O + b.G + b.A + 2);
  bool ok = cudaMalloc(&b.rec, recs * rw * 4) == cudaSuccess && cudaMalloc(&b.dg, eps * b.G * 4) == cudaSuccess &&
            cudaMalloc(&b.ep_len, eps * 4) == cudaSuccess && cudaMalloc(&b.cur_k, b.N * 4) == cudaSuccess &&
            cudaMalloc(&b.cur_t, b.N * 4) == cudaSuccess && cudaMalloc(&b.counters, 4 * sizeof(unsigned long long)) == cudaSuccess;
  if (!ok) { her_free(h); delete h; return fail(XARM_E_NOMEM, "xarm_her_create: cudaMalloc failed (N x K x (T+1) x (O + G + A + 2) floats)"); }
  // records are zero until written: a gather never reads uninitialised memory even if begin() was skipped
  cudaMemset(b.rec, 0, recs * rw * 4); cudaMemset(b.dg, 0, eps * b.G * 4);
  cudaMemset(b.ep_len, 0, eps * 4); cudaMemset(b.cur_k, 0, b.N * 4); cudaMemset(b.cur_t, 0, b.N * 4);
  cudaMemset(b.counters, 0, 4 * sizeof(unsigned long long));
  CUDA_TRY(cudaDeviceSynchronize());
  *out = h;
  return XARM_OK;
}

int xarm_her_destroy(XarmHer* h) {
  if (!h) return XARM_OK;
  cudaSetDevice(h->cfg.device);
  her_free(h);
  delete h;
  return XARM_OK;
}

int xarm_her_begin(XarmHer* h, const float* obs, const float* ag, const float* dg, const uint8_t* mask, void* stream) {
  if (!h || !obs || !ag || !dg) return fail(XARM_E_INVALID, "xarm_her_begin: null argument");
  CUDA_TRY(cudaSetDevice(h->cfg.device));
  k_her_begin<<<(unsigned)((h->b.N + 7) / 8), 256, 0, (cudaStream_t)stream>>>(h->b, obs, ag, dg, mask);   // one warp per env
  g_launches++;
  CUDA_TRY(cudaGetLastError());
  return XARM_OK;
}

int xarm_her_add(XarmHer* h, const float* obs, const float* ag, const float* dg, const float* terminal, const float* action,
                 const float* reward, const uint8_t* done, const uint8_t* truncated, void* stream) {
  if (!h || !obs || !ag || !dg || !action || !reward || !done) return fail(XARM_E_INVALID, "xarm_her_add: null argument");
  CUDA_TRY(cudaSetDevice(h->cfg.device));
  const HerBuf& b = h->b;
  cudaStream_t s = (cudaStream_t)stream;
  const unsigned grid = (unsigned)((b.N + 7) / 8);   // one warp per env
  const int nr = (b.O + b.G + b.A + 2 + 31) / 32;    // registers per lane for the run (<= 6: O <= 128, G <= 9, A <= 32)
#define HER_STORE(R) k_her_store<R><<<grid, 256, 0, s>>>(b, obs, ag, dg, terminal, action, reward, done, truncated)
  if (nr <= 1) HER_STORE(1); else if (nr == 2) HER_STORE(2); else if (nr == 3) HER_STORE(3); else if (nr == 4) HER_STORE(4); else HER_STORE(6);
#undef HER_STORE
  k_her_advance<<<(unsigned)((b.N + 255) / 256), 256, 0, s>>>(b, done);
  g_launches += 2;
  CUDA_TRY(cudaGetLastError());
  return XARM_OK;
}

int xarm_her_sample(XarmHer* h, int64_t batch, float* obs, float* ag, float* dg, float* action, float* next_obs, float* next_ag,
                    float* reward, uint8_t* done, int32_t* index, void* stream) {
  if (!h || !obs || !ag || !dg || !action || !next_obs || !next_ag || !reward || !done) return fail(XARM_E_INVALID, "xarm_her_sample: null argument");
  if (batch <= 0 || batch > 0x7fffffffLL) return fail(XARM_E_INVALID, "xarm_her_sample: 0 < batch < 2^31 required");
  CUDA_TRY(cudaSetDevice(h->cfg.device));
  cudaStream_t s = (cudaStream_t)stream;
  int4* idx = reinterpret_cast<int4*>(index);
  if (idx && (reinterpret_cast<uintptr_t>(index) & 15) != 0) return fail(XARM_E_INVALID, "xarm_her_sample: index must be 16-byte aligned");
  const HerBuf& b = h->b;
  const int64_t n_her = (int64_t)((1.0 - 1.0 / (double)(h->cfg.n_sampled_goal + 1)) * (double)batch);   // int(her_ratio * batch_size)
  const bool fused = batch <= XARM_HER_FUSED_MAX_BATCH;   // launch-bound sizes: one kernel, lane 0 of each warp draws its own index
  if (!fused && !idx) {
    if (batch > h->index_cap) {
      CUDA_TRY(cudaStreamSynchronize(s));
      cudaFree(h->index); h->index = nullptr; h->index_cap = 0;
      if (cudaMalloc(&h->index, sizeof(int4) * batch) != cudaSuccess) { cudaGetLastError(); return fail(XARM_E_NOMEM, "xarm_her_sample: cudaMalloc failed"); }
      h->index_cap = batch;
    }
    idx = h->index;
  }
  if (!fused) { k_her_index<<<(unsigned)((batch + 255) / 256), 256, 0, s>>>(b, batch, n_her, h->cfg.seed, h->calls, idx); g_launches++; }
  const unsigned grid = (unsigned)((batch + 7) / 8);   // one warp per sample
  const int nr = (2 * (b.O + b.G) + b.A + 2 + 31) / 32;   // registers per lane for the run (<= 10)
#define HER_GATHER_F(R, F) k_her_gather<R, F><<<grid, 256, 0, s>>>(b, batch, idx, n_her, h->cfg.seed, h->calls, h->cfg.task, h->cfg.reward_type, h->cfg.num_obj, obs, ag, dg, action, next_obs, next_ag, reward, done)
#define HER_GATHER(R) do { if (fused) HER_GATHER_F(R, true); else HER_GATHER_F(R, false); } while (0)
  if (nr <= 2) HER_GATHER(2); else if (nr == 3) HER_GATHER(3); else if (nr == 4) HER_GATHER(4); else if (nr == 5) HER_GATHER(5);
  else if (nr <= 7) HER_GATHER(7); else HER_GATHER(10);
#undef HER_GATHER
#undef HER_GATHER_F
  g_launches++;
  h->calls++;
  CUDA_TRY(cudaGetLastError());
  return XARM_OK;
}

int xarm_her_stats(XarmHer* h, int64_t out[4]) {
  if (!h || !out) return fail(XARM_E_INVALID, "xarm_her_stats: null argument");
  CUDA_TRY(cudaSetDevice(h->cfg.device));
  CUDA_TRY(cudaDeviceSynchronize());
  unsigned long long c[4];
  CUDA_TRY(cudaMemcpy(c, h->b.counters, sizeof(c), cudaMemcpyDeviceToHost));
  CUDA_TRY(cudaMemset(h->b.counters, 0, sizeof(unsigned long long)));   // invalid-sample counter resets on read
  out[0] = (int64_t)c[0]; out[1] = (int64_t)c[1]; out[2] = (int64_t)c[2]; out[3] = (int64_t)h->calls;
  return XARM_OK;
}

}  // extern "C"
