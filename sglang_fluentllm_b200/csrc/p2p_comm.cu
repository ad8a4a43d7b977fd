// One-box collectives over NVLink peer memory for the exchanges either side of the attention block:
//   * all-reduce of the row-parallel o_proj output (reference: tensor_model_parallel_all_reduce behind
//     python/sglang/srt/layers/linear.py:1139),
//   * uneven reduce-scatter / all-gather over tokens inside the attention-TP group -- the DP-attention conversions of
//     python/sglang/srt/distributed/decoder_comm_manager.py:42-89, whose sm_90 implementation is the triton multimem pair
//     in distributed/device_communicators/custom_triton_rsag/triton_rsag.py:97-300.
// Every rank owns one communication buffer and one flag pad, allocated here (cudaMalloc, so that a legacy CUDA IPC
// handle exists for it) and mapped into its peers (b200_comm_ipc_*).  One kernel per collective: it copies the rank's
// operand into its buffer, meets its peers, and reads their buffers with 16-byte loads through NVLink.
//   * Meeting: block b of every rank meets block b of every other rank.  A signal is a posted system-scope store of the
//     block's call counter into the peer's pad (no round trip), the wait polls the own pad; the counters live in device
//     memory and advance by two per call, so a captured graph replays without host-side state.  Every phase uses one
//     map from 16-byte units to blocks, so what block b reads from a peer is what the peer's block b wrote.
//   * The buffer is used in halves, alternating per call (parity of the call counter): a rank refills a half only two
//     calls later, after a meeting that every peer reached only once its kernel of the call before had finished
//     reading -- no barrier at the end of a call.
//   * Sums run in rank order in fp32, so every rank holds the same bits.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "b200_decode.h"
#include "host_utils.h"

namespace b200 {
namespace comm {

constexpr int kMaxWorld = 16;
constexpr int kMaxBlocks = 64;
constexpr int kMeets = 2;  // meetings per call (the second one only between the phases of the two-shot all-reduce)
constexpr int kThreads = 512;
constexpr int kPadInts = kMeets * kMaxBlocks * kMaxWorld + kMaxBlocks;  // signal slots, then one call counter per block

struct Peers {
  uint4* buf[kMaxWorld];        // every rank's communication buffer as mapped here
  int* flag[kMaxWorld];         // every rank's flag pad as mapped here: [kMeets][kMaxBlocks][kMaxWorld] + [kMaxBlocks]
  int64_t row0[kMaxWorld + 1];  // reduce-scatter / all-gather: first row of every rank's share (prefix sums)
  int64_t half_units;           // 16-byte units in one half of a communication buffer
  uint32_t spin_limit;          // polls before a meeting gives up and traps (a peer that never arrives)
};

__device__ __forceinline__ void st_release_sys(int* p, int v) {
  asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ int ld_acquire_sys(const int* p) {
  int v;
  asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// this block's call counter (the same number on every rank: the calls of a group are collective)
__device__ __forceinline__ int call_epoch(const Peers& p, int rank) {
  return *reinterpret_cast<volatile int*>(p.flag[rank] + kMeets * kMaxBlocks * kMaxWorld + blockIdx.x);
}
__device__ __forceinline__ void finish_call(const Peers& p, int rank, int epoch) {
  __syncthreads();  // every waiting thread of this block has read the counter's old value
  if (threadIdx.x == 0) p.flag[rank][kMeets * kMaxBlocks * kMaxWorld + blockIdx.x] = epoch + kMeets;
}

// Block b of this rank meets block b of every peer: what this block wrote before is visible to the peers' block b after.
__device__ __forceinline__ void meet_peers(const Peers& p, int rank, int world, int which, int epoch) {
  __syncthreads();
  if ((int)threadIdx.x < world) {
    const int t = threadIdx.x;
    const int idx = (which * kMaxBlocks + blockIdx.x) * kMaxWorld;
    const int want = epoch + which + 1;
    __threadfence_system();
    st_release_sys(p.flag[t] + idx + rank, want);
    uint32_t spins = 0;
    while (ld_acquire_sys(p.flag[rank] + idx + t) - want < 0) {  // (wrap-safe: the peers are at most one call ahead)
      if (++spins > p.spin_limit) __trap();  // a peer that never arrives: fail instead of hanging the device
    }
  }
  __syncthreads();
}

__device__ __forceinline__ void add8(float* acc, const uint4& v) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 f = __bfloat1622float2(h[i]);
    acc[2 * i] += f.x;
    acc[2 * i + 1] += f.y;
  }
}
__device__ __forceinline__ uint4 pack8(const float* acc) {
  uint4 v;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&v);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(acc[2 * i], acc[2 * i + 1]);
  return v;
}

// out[u - out0] = sum over the ranks (rank order, fp32) of unit base + u of every buffer, for this thread's units u in
// [lo, hi) (u advances by `stride`).  kUnroll units at a time with every load issued before the first add: a remote
// load is a 2-3 us round trip, so what a thread has in flight is what fills the link.
constexpr int kUnroll = 4;
template <int W>
__device__ __forceinline__ void reduce_range_w(const Peers& p, int64_t base, int64_t u, int64_t hi, int64_t stride,
                                               uint4* __restrict__ out, int64_t out0) {
  constexpr int kUnroll = W >= 8 ? 1 : 8 / W;  // eight 16-byte loads in flight per thread
  for (; u < hi; u += kUnroll * stride) {
    uint4 v[kUnroll][W];
#pragma unroll
    for (int j = 0; j < kUnroll; ++j)
      if (u + j * stride < hi) {
#pragma unroll
        for (int r = 0; r < W; ++r) v[j][r] = __ldcg(p.buf[r] + base + u + j * stride);
      }
#pragma unroll
    for (int j = 0; j < kUnroll; ++j)
      if (u + j * stride < hi) {
        float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int r = 0; r < W; ++r) add8(acc, v[j][r]);
        out[u + j * stride - out0] = pack8(acc);
      }
  }
}
__device__ __forceinline__ void reduce_range(const Peers& p, int world, int64_t base, int64_t u, int64_t hi, int64_t stride,
                                             uint4* __restrict__ out, int64_t out0) {
  switch (world) {  // the group sizes of one box; anything else takes the generic loop
    case 2: return reduce_range_w<2>(p, base, u, hi, stride, out, out0);
    case 4: return reduce_range_w<4>(p, base, u, hi, stride, out, out0);
    case 8: return reduce_range_w<8>(p, base, u, hi, stride, out, out0);
    default:
      for (; u < hi; u += stride) {
        float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        for (int r = 0; r < world; ++r) add8(acc, __ldcg(p.buf[r] + base + u));
        out[u - out0] = pack8(acc);
      }
  }
}
// out[u] = src[u] for this thread's units in [lo, hi), kUnroll loads in flight
__device__ __forceinline__ void copy_range(const uint4* __restrict__ src, uint4* __restrict__ dst, int64_t u, int64_t hi,
                                           int64_t stride) {
  for (; u < hi; u += kUnroll * stride) {
    uint4 v[kUnroll];
#pragma unroll
    for (int j = 0; j < kUnroll; ++j)
      if (u + j * stride < hi) v[j] = __ldcg(src + u + j * stride);
#pragma unroll
    for (int j = 0; j < kUnroll; ++j)
      if (u + j * stride < hi) dst[u + j * stride] = v[j];
  }
}

// The one map from units to threads that every phase uses: unit u belongs to thread (u mod stride).  First unit of
// this thread inside [lo, hi).
__device__ __forceinline__ int64_t first_unit(int64_t lo) {
  const int64_t stride = (int64_t)gridDim.x * kThreads;
  const int64_t mine = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  return lo + ((mine - lo) % stride + stride) % stride;
}

// All-reduce.  Small groups (one_shot): every rank sums every unit of every buffer -- (w - 1) messages over NVLink per
// rank, one meeting.  Larger groups: rank r sums slice r into its own buffer, a second meeting, everybody collects the
// reduced slices -- 2 (w - 1) / w messages.  slice r = units [r * per, min((r + 1) * per, units)).
__global__ void __launch_bounds__(kThreads) all_reduce_kernel(const __grid_constant__ Peers p, int rank, int world,
                                                              const uint4* __restrict__ in, int64_t units,
                                                              uint4* __restrict__ out, int one_shot) {
  const int epoch = call_epoch(p, rank);
  const int64_t base = ((epoch / kMeets) & 1) * p.half_units;
  const int64_t stride = (int64_t)gridDim.x * kThreads;
  uint4* mine = p.buf[rank] + base;
  copy_range(in, mine, first_unit(0), units, stride);
  meet_peers(p, rank, world, 0, epoch);  // every rank's operand is in its buffer
  if (one_shot) {
    reduce_range(p, world, base, first_unit(0), units, stride, out, 0);
  } else {
    const int64_t per = (units + world - 1) / world;
    {
      const int64_t lo = rank * per, hi = min(lo + per, units);
      reduce_range(p, world, base, first_unit(lo), hi, stride, mine, 0);
    }
    meet_peers(p, rank, world, 1, epoch);
    for (int r = 0; r < world; ++r) {
      const int64_t lo = r * per, hi = min(lo + per, units);
      copy_range(p.buf[r] + base, out, first_unit(lo), hi, stride);
    }
  }
  finish_call(p, rank, epoch);
}

// All-reduce through the switch (NVLS): `mc` is a multicast address that maps the same symmetric buffer on every rank.
// Rank r asks the switch for the sum of slice r over all ranks (multimem.ld_reduce: one response per 16 bytes instead of
// one per rank) and broadcasts it back into every rank's buffer (multimem.st); after a second meeting every rank copies
// the result out of its own buffer.  Per rank 2 / w of the message crosses its link.
__device__ __forceinline__ uint4 multimem_ld_reduce_bf16(const uint4* mc) {
  uint4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.acc::f32.v4.bf16x2 {%0, %1, %2, %3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "l"(mc)
               : "memory");
  return v;
}
__device__ __forceinline__ void multimem_st_bf16(uint4* mc, const uint4& v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.bf16x2 [%0], {%1, %2, %3, %4};" ::"l"(mc), "r"(v.x), "r"(v.y), "r"(v.z),
               "r"(v.w)
               : "memory");
}

__global__ void __launch_bounds__(kThreads) all_reduce_nvls_kernel(const __grid_constant__ Peers p, int rank, int world,
                                                                   uint4* __restrict__ mc, uint4* __restrict__ local,
                                                                   const uint4* __restrict__ in, int64_t units,
                                                                   uint4* __restrict__ out) {
  const int epoch = call_epoch(p, rank);
  const int64_t base = ((epoch / kMeets) & 1) * p.half_units;
  const int64_t stride = (int64_t)gridDim.x * kThreads;
  copy_range(in, local + base, first_unit(0), units, stride);
  meet_peers(p, rank, world, 0, epoch);
  {
    const int64_t per = (units + world - 1) / world;
    const int64_t lo = rank * per, hi = min(lo + per, units);
    for (int64_t u = first_unit(lo); u < hi; u += kUnroll * stride) {
      uint4 v[kUnroll];
#pragma unroll
      for (int j = 0; j < kUnroll; ++j)
        if (u + j * stride < hi) v[j] = multimem_ld_reduce_bf16(mc + base + u + j * stride);
#pragma unroll
      for (int j = 0; j < kUnroll; ++j)
        if (u + j * stride < hi) multimem_st_bf16(mc + base + u + j * stride, v[j]);
    }
  }
  meet_peers(p, rank, world, 1, epoch);
  copy_range(local + base, out, first_unit(0), units, stride);
  finish_call(p, rank, epoch);
}

// Uneven reduce-scatter over rows: out[rows of this rank] = sum over ranks of their operands' same rows.
__global__ void __launch_bounds__(kThreads) reduce_scatter_kernel(const __grid_constant__ Peers p, int rank, int world,
                                                                  const uint4* __restrict__ in, int64_t units_per_row,
                                                                  uint4* __restrict__ out) {
  const int epoch = call_epoch(p, rank);
  const int64_t base = ((epoch / kMeets) & 1) * p.half_units;
  const int64_t stride = (int64_t)gridDim.x * kThreads;
  const int64_t units = p.row0[world] * units_per_row;
  uint4* mine = p.buf[rank] + base;
  copy_range(in, mine, first_unit(0), units, stride);
  meet_peers(p, rank, world, 0, epoch);
  const int64_t lo = p.row0[rank] * units_per_row, hi = p.row0[rank + 1] * units_per_row;
  reduce_range(p, world, base, first_unit(lo), hi, stride, out, lo);
  finish_call(p, rank, epoch);
}

// Uneven all-gather over rows: every rank puts its rows at their global position in its own buffer.
__global__ void __launch_bounds__(kThreads) all_gather_kernel(const __grid_constant__ Peers p, int rank, int world,
                                                              const uint4* __restrict__ in, int64_t units_per_row,
                                                              uint4* __restrict__ out) {
  const int epoch = call_epoch(p, rank);
  const int64_t base = ((epoch / kMeets) & 1) * p.half_units;
  const int64_t stride = (int64_t)gridDim.x * kThreads;
  {
    const int64_t lo = p.row0[rank] * units_per_row, hi = p.row0[rank + 1] * units_per_row;
    uint4* mine = p.buf[rank] + base;
    for (int64_t u = first_unit(lo); u < hi; u += stride) mine[u] = in[u - lo];
  }
  meet_peers(p, rank, world, 0, epoch);
  for (int r = 0; r < world; ++r) {
    const int64_t lo = p.row0[r] * units_per_row, hi = p.row0[r + 1] * units_per_row;
    copy_range(r == rank ? in - lo : p.buf[r] + base, out, first_unit(lo), hi, stride);
  }
  finish_call(p, rank, epoch);
}

static int fill_peers(Peers& p, void* const* bufs, void* const* flags, size_t buf_bytes, int rank, int world,
                      const int64_t* row_offsets) {
  B200_CHECK(bufs && flags, "comm: null pointer table");
  p.half_units = (int64_t)(buf_bytes / 32);
  B200_CHECK(world >= 1 && world <= kMaxWorld && rank >= 0 && rank < world, "comm: rank %d of %d (at most %d ranks)", rank,
             world, kMaxWorld);
  for (int r = 0; r < kMaxWorld; ++r) {
    p.buf[r] = r < world ? (uint4*)bufs[r] : nullptr;
    p.flag[r] = r < world ? (int*)flags[r] : nullptr;
    B200_CHECK(r >= world || (bufs[r] && flags[r] && ((uintptr_t)bufs[r] & 15) == 0), "comm: buffer of rank %d missing or unaligned", r);
  }
  for (int r = 0; r <= kMaxWorld; ++r) p.row0[r] = row_offsets && r <= world ? row_offsets[r] : 0;
  // about a minute of polling by default: ranks of a serving process can be seconds apart at their first exchange
  static const uint32_t limit = getenv("B200_COMM_SPIN_LIMIT") ? (uint32_t)atoll(getenv("B200_COMM_SPIN_LIMIT")) : (1u << 26);
  p.spin_limit = limit;
  return 0;
}

// Every call of a process uses the same grid: the halves of the buffers alternate with the blocks' call counters, which
// must therefore advance together (B200_COMM_BLOCKS, the same on every rank, overrides the default).
static int pick_blocks(int64_t, int) {
  static const int n = getenv("B200_COMM_BLOCKS") ? atoi(getenv("B200_COMM_BLOCKS")) : 48;
  return n < 1 ? 1 : (n > kMaxBlocks ? kMaxBlocks : n);
}

}  // namespace comm
}  // namespace b200

using namespace b200;
using namespace b200::comm;

extern "C" {

size_t b200_comm_flag_bytes(void) { return sizeof(int) * kPadInts; }

int b200_comm_alloc(size_t bytes, void** ptr) {
  B200_CHECK(ptr && bytes > 0, "comm_alloc: bad arguments");
  B200_CUDA(cudaMalloc(ptr, bytes));
  B200_CUDA(cudaMemset(*ptr, 0, bytes));
  B200_CUDA(cudaDeviceSynchronize());
  return 0;
}

int b200_comm_free(void* ptr) {
  if (ptr) B200_CUDA(cudaFree(ptr));
  return 0;
}

int b200_comm_ipc_export(void* ptr, void* handle64) {
  B200_CHECK(ptr && handle64, "comm_ipc_export: null pointer");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle size");
  B200_CUDA(cudaIpcGetMemHandle(reinterpret_cast<cudaIpcMemHandle_t*>(handle64), ptr));
  return 0;
}

int b200_comm_ipc_open(const void* handle64, void** ptr) {
  B200_CHECK(ptr && handle64, "comm_ipc_open: null pointer");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, sizeof(h));
  B200_CUDA(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return 0;
}

int b200_comm_ipc_close(void* ptr) {
  if (ptr) B200_CUDA(cudaIpcCloseMemHandle(ptr));
  return 0;
}

int b200_comm_all_reduce_bf16(void* const* bufs, void* const* flags, size_t buf_bytes, int rank, int world, const void* in,
                              int64_t numel, void* out, int blocks, void* stream) {
  Peers p;
  if (fill_peers(p, bufs, flags, buf_bytes, rank, world, nullptr)) return 1;
  B200_CHECK(in && out && numel > 0 && numel % 8 == 0 && ((uintptr_t)out & 15) == 0 && ((uintptr_t)in & 15) == 0,
             "comm_all_reduce: numel %lld must be a positive multiple of 8, in / out 16-byte aligned", (long long)numel);
  const int64_t units = numel / 8;
  B200_CHECK(units <= p.half_units, "comm_all_reduce: %lld bytes do not fit half of the %zu-byte communication buffer",
             (long long)numel * 2, buf_bytes);
  // two ranks: one pass over the peer's operand; more: reduce a slice each, then collect
  static const int force = getenv("B200_COMM_ONE_SHOT") ? atoi(getenv("B200_COMM_ONE_SHOT")) : -1;
  // measured on 8 GPUs: one pass 36 us, two-shot 41-47 us for 1.8 MB (a meeting costs more than the extra traffic)
  const int one_shot = force >= 0 ? force : ((int64_t)(world - 1) * numel * 2 <= (16ll << 20));
  all_reduce_kernel<<<pick_blocks(one_shot ? units : (units + world - 1) / world, blocks), kThreads, 0, (cudaStream_t)stream>>>(
      p, rank, world, (const uint4*)in, units, (uint4*)out, one_shot);
  B200_CUDA(cudaGetLastError());
  return 0;
}

int b200_comm_all_reduce_nvls_bf16(void* multicast_ptr, void* local_ptr, void* const* flags, size_t buf_bytes, int rank,
                                   int world, const void* in, int64_t numel, void* out, int blocks, void* stream) {
  Peers p;
  void* none[kMaxWorld] = {};
  for (int r = 0; r < world && r < kMaxWorld; ++r) none[r] = local_ptr;  // (the peers' buffers are reached through the switch)
  if (fill_peers(p, none, flags, buf_bytes, rank, world, nullptr)) return 1;
  B200_CHECK(multicast_ptr && local_ptr && ((uintptr_t)multicast_ptr & 15) == 0, "comm_all_reduce_nvls: no multicast mapping");
  B200_CHECK(in && out && numel > 0 && numel % 8 == 0 && ((uintptr_t)out & 15) == 0 && ((uintptr_t)in & 15) == 0,
             "comm_all_reduce_nvls: numel %lld must be a positive multiple of 8, in / out 16-byte aligned", (long long)numel);
  const int64_t units = numel / 8;
  B200_CHECK(units <= p.half_units, "comm_all_reduce_nvls: %lld bytes do not fit half of the %zu-byte symmetric buffer",
             (long long)numel * 2, buf_bytes);
  all_reduce_nvls_kernel<<<pick_blocks(units, blocks), kThreads, 0, (cudaStream_t)stream>>>(
      p, rank, world, (uint4*)multicast_ptr, (uint4*)local_ptr, (const uint4*)in, units, (uint4*)out);
  B200_CUDA(cudaGetLastError());
  return 0;
}

int b200_comm_reduce_scatter_bf16(void* const* bufs, void* const* flags, size_t buf_bytes, int rank, int world,
                                  const int64_t* row_offsets, const void* in, int64_t hidden, void* out, int blocks,
                                  void* stream) {
  Peers p;
  B200_CHECK(row_offsets, "comm_reduce_scatter: null row offsets");
  if (fill_peers(p, bufs, flags, buf_bytes, rank, world, row_offsets)) return 1;
  B200_CHECK(hidden > 0 && hidden % 8 == 0, "comm_reduce_scatter: hidden %lld must be a multiple of 8", (long long)hidden);
  const int64_t rows = row_offsets[rank + 1] - row_offsets[rank], total = row_offsets[world];
  B200_CHECK(rows >= 0 && (rows == 0 || (out && ((uintptr_t)out & 15) == 0)), "comm_reduce_scatter: bad output");
  B200_CHECK(total == 0 || (in && ((uintptr_t)in & 15) == 0), "comm_reduce_scatter: bad input");
  B200_CHECK(total * (hidden / 8) <= p.half_units, "comm_reduce_scatter: %lld rows do not fit half of the communication buffer",
             (long long)total);
  // every rank launches (its peers wait for it) with the same number of blocks
  reduce_scatter_kernel<<<pick_blocks(total * (hidden / 8), blocks), kThreads, 0, (cudaStream_t)stream>>>(
      p, rank, world, (const uint4*)in, hidden / 8, (uint4*)out);
  B200_CUDA(cudaGetLastError());
  return 0;
}

int b200_comm_all_gather_bf16(void* const* bufs, void* const* flags, size_t buf_bytes, int rank, int world,
                              const int64_t* row_offsets, const void* in, int64_t hidden, void* out, int blocks,
                              void* stream) {
  Peers p;
  B200_CHECK(row_offsets, "comm_all_gather: null row offsets");
  if (fill_peers(p, bufs, flags, buf_bytes, rank, world, row_offsets)) return 1;
  B200_CHECK(hidden > 0 && hidden % 8 == 0, "comm_all_gather: hidden %lld must be a multiple of 8", (long long)hidden);
  const int64_t rows = row_offsets[rank + 1] - row_offsets[rank], total = row_offsets[world];
  B200_CHECK(total == 0 || (out && ((uintptr_t)out & 15) == 0), "comm_all_gather: bad output");
  B200_CHECK(rows == 0 || (in && ((uintptr_t)in & 15) == 0), "comm_all_gather: bad input");
  B200_CHECK(total * (hidden / 8) <= p.half_units, "comm_all_gather: %lld rows do not fit half of the communication buffer",
             (long long)total);
  all_gather_kernel<<<pick_blocks(total * (hidden / 8), blocks), kThreads, 0, (cudaStream_t)stream>>>(
      p, rank, world, (const uint4*)in, hidden / 8, (uint4*)out);
  B200_CUDA(cudaGetLastError());
  return 0;
}

}  // extern "C"
