// Label fusion for sm_100a: MeshAggregator.add / get (include/semantic_meshes/fusion/Mesh.h:57-133 with the aggregator
// chains of python/semantic_meshes/src/Fusion.cu:46-92).
//
// The reference copies the view to the host, builds a serial std::map histogram and then does a mutex-guarded vector add
// per pixel under OpenMP. Here one view is three launches on the caller's stream:
//   1. count_runs_kernel - per-face pixel count of this view (Mesh.h:90-93), runs of equal ids inside a warp merged into
//                       one atomicAdd; also writes the flat-order uint32 copy of the ids when the input is strided or
//                       not 32-bit
//   2. scatter_kernel - THE hot kernel, HBM-bound: the (n_pix, C) probability image is streamed exactly once through a
//                       multi-stage shared-memory ring filled by TMA bulk copies (cp.async.bulk + mbarrier, evict-first
//                       L2 policy so the stream does not flush the accumulator rows / ids / counters out of L2); each
//                       consumer warp owns 32 consecutive pixels of a stage: lane = pixel for the gate (sequential class
//                       sum, Mesh.h:98) and weight (Mesh.h:100-103), then lanes regroup as (run of equal face id, 4-class
//                       chunk) to reduce the run in registers and issue ONE 128-bit red.global.add.v4.f32 per chunk into
//                       the 16-byte padded accumulator row
//                       (C = 2 ... 20: scatter_pair_kernel, two pixels per lane; wide C: scatter_rows_kernel, lanes across the
//                       classes of a pixel, rows read straight from global memory)
//   3. clear_kernel   - only in the untagged-counter mode (images of >= 2^24 pixels): zero the touched counters again
// A batch (smesh_fuse_add_batch) deals its views to two lanes - the caller's stream and a side stream of the library, one
// counter array each - so one view's stage 1, launch gaps and tail run under another view's stage 2. For callers that
// hold the next index image early (smesh_fuse_scatter_count_next; also the batch when it has to stay on one stream)
// stage 1 of the NEXT view rides in stage 2 of this one: one more warp per CTA of the ring kernels (count_job_warp)
// takes the next view's counts into the other of two counter arrays.
// get() (get_stream_kernel): every warp runs a two-stage bulk-copy pipeline over blocks of 32 accumulator rows - bulk
// load, one lane per row, bulk store (get_kernel: the staged variant for very wide class vectors).
// No tensor cores: this is an irregular gather/scatter, not a contraction.
#include "smesh_common.cuh"

#include <math_constants.h>
#include <stdlib.h>

#include <algorithm>

namespace smesh {
namespace fuse {

constexpr uint32_t INVALID_ID = 0xFFFFFFFFu;

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d)
{
  asm volatile("red.relaxed.gpu.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

__device__ __forceinline__ void red_add_f32(float* addr, float a)
{
  asm volatile("red.relaxed.gpu.global.add.f32 [%0], %1;" ::"l"(addr), "f"(a) : "memory");
}

// ---------------------------------------------------------------------------------------------------------------------
// shared arithmetic
// ---------------------------------------------------------------------------------------------------------------------

// Mesh.h:100-103: image_pixel_weight = iew * (1 / n) + (1 - iew) * 1, times the pixel's weight. Plain IEEE float ops,
// no contraction (the reference host build has none either).
__device__ __forceinline__ float pixel_weight(float iew, uint32_t n, float wt)
{
  const float image_weight = __fdiv_rn(1.0f, (float) n);
  const float w = __fadd_rn(__fmul_rn(iew, image_weight), __fmul_rn(__fsub_rn(1.0f, iew), 1.0f));
  return __fmul_rn(w, wt);
}

// log(x) for a NORMAL, positive, finite x (the caller checks): exponent / mantissa split with the mantissa in [2/3, 4/3),
// log1p(f) = f - f^2/2 + f^3 Q(f) with a degree-7 Q fitted over [-1/3, 1/3]. Maximum error 0.77 ulp on the mantissa part
// (all 2^23 x 2/3 floats of [2/3, 4/3) checked, tests/golden/fit_log.py), ~1.3 ulp overall: the same class as logf (1 ulp),
// at 17 instructions without logf's handling of zero, denormals, infinities and NaN.
__device__ __forceinline__ float log_normal(float x)
{
  const int i = __float_as_int(x);
  const int e = (i - 0x3f2aaaab) & (int) 0xff800000;
  const float f = __int_as_float(i - e) - 1.0f;
  const float fe = (float) (e >> 23);
  float q = -0.12890197336673737f;
  q = fmaf(q, f, 0.13985218107700348f);
  q = fmaf(q, f, -0.12184639275074005f);
  q = fmaf(q, f, 0.14005699753761292f);
  q = fmaf(q, f, -0.16680456697940826f);
  q = fmaf(q, f, 0.20010416209697723f);
  q = fmaf(q, f, -0.24999798834323883f);
  q = fmaf(q, f, 0.3333321511745453f);
  const float s = f * f;
  const float r = fmaf(fmaf(q, f, -0.5f), s, f);
  return fmaf(fe, 0.693147182f, r);
}

// mul aggregator input: LogProb(pow(p, w)) as -log (Fusion.cu:83-87, tt/numeric/LogProb.h:66-71); "zero" (isinf of
// either sign) is the absorbing +inf (LogProb.h:106-118).
// The reference's powf + logf pair costs ~150 instructions per class. -log(p^w) = -w log p, and as long as q = p^w stays a
// normal float (|w log p| < 80, p a normal positive float, w > 0) the reference's value is -log(fl(q)) = -w log p within
// 6e-8 absolute (the rounding of q, which the direct form does not even have) + 2e-7 relative (the two logarithms): the
// direct form is used there, with log_normal() above. Anything else - underflow of q to the absorbing zero, overflow,
// p <= 0 or denormal, NaN, w <= 0 - takes the reference's own sequence. `exact` (SMESH_MUL_EXACT=1, verification) forces
// that sequence everywhere.
// (out of line: inlined at each of the 38 call sites of a tile, powf + logf made the loop body of the mul kernels some
// 4000 instructions long - more than the instruction caches hold)
__device__ __noinline__ float neg_log_pow_reference(float p, float w)
{
  const float q = powf(p, w);
  float l = (q == 0.0f) ? CUDART_INF_F : -logf(q);
  if (isinf(l))
  {
    l = CUDART_INF_F;
  }
  return l;
}

// the direct form of one element, computed unconditionally; `bad` is raised where it does not apply to the element
__device__ __forceinline__ float neg_log_pow_direct(float p, float w, bool& bad)
{
  const float l = -w * log_normal(p);
  bad |= !(__float_as_uint(p) - 0x00800000u < 0x7F000000u) || !(fabsf(l) < 80.0f); // 0x00800000 <= bits < 0x7F800000: normal, positive
  return l;
}

// one element: the direct form where it applies, else the reference's sequence
__device__ __forceinline__ float neg_log_pow(float p, float w, bool exact)
{
  bool bad = exact || !(w > 0.0f);
  const float l = neg_log_pow_direct(p, w, bad);
  return bad ? neg_log_pow_reference(p, w) : l;
}

// (the per-element choice out of line, for the rare second pass of a class vector that held an element the direct form
// does not cover: the kernels evaluate a whole vector in the direct form without branches first)
__device__ __noinline__ float neg_log_pow_checked(float p, float w, bool exact)
{
  return neg_log_pow(p, w, exact);
}

// ---------------------------------------------------------------------------------------------------------------------
// 1. / 3. per-face pixel count of one view and its reset
//
// counts[id] is a tagged word: (epoch << 24) | n. A view with epoch e first raises the word to at least e << 24
// (atomicMax) and then adds its pixels, so words left over from earlier views (smaller epoch) never need clearing; the
// caller zero-fills the array only when the 8-bit epoch wraps. epoch 0 = untagged 32-bit counts + clear_kernel afterwards
// (images with >= 2^24 pixels).
// ---------------------------------------------------------------------------------------------------------------------

constexpr uint32_t COUNT_BITS = 24;
constexpr uint32_t COUNT_MASK = (1u << COUNT_BITS) - 1u;

// Mesh.h:95 `primitive_index < rows()` on a size_t: negative values wrap to huge numbers and fail the test.
__device__ __forceinline__ uint32_t sanitize_id(uint32_t raw, int64_t P)
{
  return (int64_t) raw < P ? raw : INVALID_ID;
}
__device__ __forceinline__ uint32_t sanitize_id(int32_t raw, int64_t P)
{
  return (raw >= 0 && (int64_t) raw < P) ? (uint32_t) raw : INVALID_ID;
}
__device__ __forceinline__ uint32_t sanitize_id(uint64_t raw, int64_t P)
{
  return raw < (uint64_t) P ? (uint32_t) raw : INVALID_ID;
}
__device__ __forceinline__ uint32_t sanitize_id(int64_t raw, int64_t P)
{
  return (raw >= 0 && raw < P) ? (uint32_t) raw : INVALID_ID;
}

// Flat order, 4 independent 32-pixel groups per warp iteration, runs of equal ids inside a group merged by ballot into one
// pair of reductions. The stage is bound by launch + load latency, not by its reductions (1.38 M per cfg3 view): without
// any reduction it still takes 5.4 of its 8.3 us. Folding the runs of a face across 2 / 4 / 8 adjacent columns in
// registers halved the L2 reduction sectors but doubled the instructions and was slower (10 - 11 us; round 2,
// profiles/r02b_count_variants.txt), like round 1's per-CTA shared-memory hash (12 - 14 us).
constexpr int COUNT_UNROLL = 4; // independent 32-pixel groups per warp iteration (loads in flight per lane)

template <typename IdT, int UNROLL>
__global__ void __launch_bounds__(256) count_runs_kernel(const IdT* __restrict__ ids, int64_t stride_outer, int64_t stride_inner,
                                                         int64_t n_inner, int64_t npix, int64_t P, uint32_t* __restrict__ counts,
                                                         uint32_t* __restrict__ ids32, int flat, uint32_t epoch)
{
  const int lane = threadIdx.x & 31;
  const int64_t warp_global = ((int64_t) blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t) gridDim.x * blockDim.x) >> 5;
  const uint32_t tag = epoch << COUNT_BITS;
  for (int64_t base = warp_global * (32 * UNROLL); base < npix; base += nwarps * (32 * UNROLL))
  {
    uint32_t id[UNROLL];
#pragma unroll
    for (int k = 0; k < UNROLL; k++)
    {
      const int64_t i = base + k * 32 + lane;
      id[k] = INVALID_ID;
      if (i < npix)
      {
        int64_t off = i;
        if (!flat)
        {
          const int64_t o = i / n_inner, in = i - o * n_inner;
          off = o * stride_outer + in * stride_inner;
        }
        id[k] = sanitize_id(ids[off], P);
      }
    }
#pragma unroll
    for (int k = 0; k < UNROLL; k++)
    {
      const int64_t i = base + k * 32 + lane;
      if (ids32 != nullptr && i < npix)
      {
        ids32[i] = id[k];
      }
      // one atomic per run of equal ids inside the 32-pixel group
      const uint32_t prev = __shfl_up_sync(0xFFFFFFFFu, id[k], 1);
      const bool head = (lane == 0) || (prev != id[k]);
      const uint32_t headmask = __ballot_sync(0xFFFFFFFFu, head);
      if (head && id[k] != INVALID_ID)
      {
        const uint32_t above = headmask & ~((2u << lane) - 1u);
        const int next = above ? (__ffs(above) - 1) : 32;
        if (epoch != 0)
        {
          atomicMax(counts + id[k], tag);
        }
        atomicAdd(counts + id[k], (uint32_t) (next - lane));
      }
    }
  }
}

// The count stage (flat uint32 ids, tagged counters) as a job of ONE warp of a persistent CTA: the CTA's share of the
// image in chunks of COUNT_UNROLL x 32 pixels, runs inside a 32-pixel group merged into one pair of reductions.
__device__ __forceinline__ void count_job_warp(const uint32_t* __restrict__ ids, int64_t npix, uint32_t P32,
                                               uint32_t* __restrict__ counts, uint32_t tag, int lane)
{
  for (int64_t base = (int64_t) blockIdx.x * (32 * COUNT_UNROLL); base < npix; base += (int64_t) gridDim.x * (32 * COUNT_UNROLL))
  {
    uint32_t id[COUNT_UNROLL];
#pragma unroll
    for (int k = 0; k < COUNT_UNROLL; k++)
    {
      const int64_t i = base + k * 32 + lane;
      id[k] = i < npix ? __ldg(ids + i) : INVALID_ID;
      if (!(id[k] < P32))
      {
        id[k] = INVALID_ID;
      }
    }
#pragma unroll
    for (int k = 0; k < COUNT_UNROLL; k++)
    {
      const uint32_t prev = __shfl_up_sync(0xFFFFFFFFu, id[k], 1);
      const bool head = (lane == 0) || (prev != id[k]);
      const uint32_t headmask = __ballot_sync(0xFFFFFFFFu, head);
      if (head && id[k] != INVALID_ID)
      {
        const uint32_t above = headmask & ~((2u << lane) - 1u);
        const int next = above ? (__ffs(above) - 1) : 32;
        atomicMax(counts + id[k], tag);
        atomicAdd(counts + id[k], (uint32_t) (next - lane));
      }
    }
  }
}

// (out of line for scatter_kernel: inlined, the job changed the register allocation of the consumers' loop and the C = 40
// kernel lost a quarter of its speed)
__device__ __noinline__ void count_job_warp_call(const uint32_t* ids, int64_t npix, uint32_t P32, uint32_t* counts, uint32_t tag,
                                                 int lane)
{
  count_job_warp(ids, npix, P32, counts, tag, lane);
}

__global__ void __launch_bounds__(256) clear_kernel(const uint32_t* __restrict__ ids32, int64_t npix, int64_t P,
                                                    uint32_t* __restrict__ counts)
{
  const int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (i < npix)
  {
    const uint32_t id = ids32[i];
    if ((int64_t) id < P && (i == 0 || ids32[i - 1] != id))
    {
      counts[id] = 0;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// 2. scatter
// ---------------------------------------------------------------------------------------------------------------------

struct ScatterArgs
{
  const float* probs;      // [npix][C]
  const uint32_t* ids;     // [npix] flat order; anything >= P is background
  const float* weights;    // NULL or [npix]
  const uint32_t* counts;  // [P] tagged pixel counts of this view
  float* acc;              // [P][Cpad]
  int64_t npix;
  int64_t P;
  int64_t ntiles;
  int C, Cpad;
  int stages;
  uint32_t count_mask;     // COUNT_MASK with a tagged epoch, 0xFFFFFFFF without
  uint32_t run_cap;        // power of two <= 32: runs of equal ids are cut every run_cap lanes
  int mul_exact;           // mul: the reference's powf + logf sequence for every element (see neg_log_pow)
  uint32_t wait_hint;      // suspend-time hint of the ring's mbarrier waits (see mbar_wait)
  // Optional second job of the ring kernels: the count stage of the NEXT view of a batch (smesh_fuse_add_batch), done by
  // one extra warp per CTA while the consumer warps scatter this view. NULL = none.
  const uint32_t* next_ids;   // [next_npix] flat uint32 ids of the next view
  uint32_t* next_counts;      // [P] the OTHER counter array
  int64_t next_npix;
  uint32_t next_tag;          // its epoch << COUNT_BITS (tagged counters only)

  float iew;
};

constexpr int CH = 20; // classes held in registers per pass (multiple of 4); C = 19 is one pass

// One pixel's classes [c0, c0 + CH) from shared memory into registers, zero beyond C. `al` = largest power of two
// <= 4 dividing C (and c0): rows of a C % 4 == 0 image are 16-byte aligned and read with 128-bit loads (a lane stride of
// C words would otherwise be an 8-way bank conflict at C = 40), C % 2 == 0 with 64-bit loads; odd C is conflict-free.
__device__ __forceinline__ void load_chunk(const float* __restrict__ p, int nvalid, int al, float (&v)[CH])
{
  if (al == 4 && nvalid == CH)
  {
#pragma unroll
    for (int j = 0; j < CH / 4; j++)
    {
      const float4 t = *reinterpret_cast<const float4*>(p + 4 * j);
      v[4 * j + 0] = t.x; v[4 * j + 1] = t.y; v[4 * j + 2] = t.z; v[4 * j + 3] = t.w;
    }
  }
  else if (al >= 2 && nvalid == CH)
  {
#pragma unroll
    for (int j = 0; j < CH / 2; j++)
    {
      const float2 t = *reinterpret_cast<const float2*>(p + 2 * j);
      v[2 * j + 0] = t.x; v[2 * j + 1] = t.y;
    }
  }
  else
  {
#pragma unroll
    for (int k = 0; k < CH; k++)
    {
      v[k] = (k < nvalid) ? p[k] : 0.0f;
    }
  }
}

// Shared memory: [stages][NW*32*C] floats | full[stages], empty[stages] mbarriers
// The 128-bit row loads of the consumers (lane stride C words) are free of bank conflicts only when C / 4 is odd. For
// C / 4 = 2 (mod 4) - C = 24, 40, 56 - lanes l and l + 4 of a quarter warp meet in the same banks, for C / 4 = 4 (mod 8) -
// C = 48 - lanes l and l + 2: PADG = 4 / 2 shifts every group of PADG pixels by one more 16-byte chunk in shared memory
// (the tile then arrives as one bulk copy per group, issued by the lanes of the producer warp in parallel).
// Only where a group is a bulk copy of at least 512 bytes (C = 40: 640, C = 56: 896): many smaller copies cost more than
// the conflicts they remove (measured on the pair kernel, profiles/r02s_pair_even_class_counts.txt). C = 40 (cfg2):
// 22.2 -> 20.1 us per view.
__host__ __device__ constexpr int ring_pad_group(int C)
{
  const int g = (C % 8 != 0) ? 0 : ((C / 4) % 4 == 2 ? 4 : ((C / 4) % 8 == 4 ? 2 : 0));
  return g * C * 4 >= 512 ? g : 0;
}

// RIDER: the CTA has one more warp, which runs the next view's count stage (see ScatterArgs)
template <int KIND, int CT, bool RIDER, int PADG>
__global__ void __launch_bounds__(RIDER ? 320 : 288, RIDER ? 3 : 0) scatter_kernel(ScatterArgs a)
{
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int C = CT > 0 ? CT : a.C;
  const int Cpad = CT > 0 ? ((CT + 3) & ~3) : a.Cpad;
  const int al = (C % 4 == 0) ? 4 : ((C % 2 == 0) ? 2 : 1);
  const int NW = (int) (blockDim.x >> 5) - 1 - (RIDER ? 1 : 0); // consumer warps; warp 0 produces, the last one may count
  const int tile_px = NW * 32;
  const size_t stage_floats = (size_t) tile_px * C + (PADG > 0 ? (size_t) (tile_px / PADG) * 4 : 0);
  const int stages = a.stages;

  float* stage_base = reinterpret_cast<float*>(smem_raw);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(stage_base + stage_floats * stages);
  uint64_t* empty_bar = full_bar + stages;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0)
  {
    for (int s = 0; s < stages; s++)
    {
      mbar_init(full_bar + s, 1);
      mbar_init(empty_bar + s, (uint32_t) NW);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (warp == 0 && PADG > 0)
  {
    // ===== producer, padded layout: one bulk copy per group of PADG pixels, the lanes issue them in parallel =====
    const uint64_t policy = l2_evict_first_policy();
    int s = 0;
    uint32_t use_parity = 1;
    bool first_pass = true;
    constexpr int PG = PADG > 0 ? PADG : 1;
    for (int64_t tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x)
    {
      if (!first_pass)
      {
        if (lane == 0)
        {
          mbar_wait(empty_bar + s, use_parity, a.wait_hint);
        }
        __syncwarp();
      }
      const int64_t px0 = tile * tile_px;
      const int px_n = (int) min((int64_t) tile_px, a.npix - px0);
      float* dst = stage_base + stage_floats * s;
      const float* src = a.probs + (size_t) px0 * C;
      if (lane == 0)
      {
        mbar_arrive_expect_tx(full_bar + s, (uint32_t) ((size_t) px_n * C * 4)); // C % 8 == 0: whole 16-byte chunks
      }
      __syncwarp();
      const int ngroups = (px_n + PG - 1) / PG;
      for (int g = lane; g < ngroups; g += 32)
      {
        const int gp = min(PG, px_n - g * PG);
        bulk_g2s(dst + (size_t) g * (PG * C + 4), src + (size_t) g * PG * C, (uint32_t) (gp * C * 4), full_bar + s, policy);
      }
      if (++s == stages)
      {
        s = 0;
        first_pass = false;
        use_parity ^= 1u;
      }
    }
    return;
  }
  if (warp == 0)
  {
    // ===== producer: one lane streams the tiles of this CTA into the ring =====
    if (lane == 0)
    {
      const uint64_t policy = l2_evict_first_policy();
      int s = 0;
      uint32_t use_parity = 1; // parity of the PREVIOUS use of stage s; the first pass over the ring does not wait
      bool first_pass = true;
      for (int64_t tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x)
      {
        if (!first_pass)
        {
          mbar_wait(empty_bar + s, use_parity, a.wait_hint);
        }
        const int64_t px0 = tile * tile_px;
        const int64_t px_n = min((int64_t) tile_px, a.npix - px0);
        const size_t nfloats = (size_t) px_n * C;
        const uint32_t bulk_bytes = (uint32_t) ((nfloats * 4) & ~(size_t) 15);
        float* dst = stage_base + stage_floats * s;
        const float* src = a.probs + (size_t) px0 * C;
        // the < 16-byte remainder of the very last tile cannot go through the bulk copy
        for (size_t k = bulk_bytes / 4; k < nfloats; k++)
        {
          dst[k] = src[k];
        }
        mbar_arrive_expect_tx(full_bar + s, bulk_bytes);
        if (bulk_bytes > 0)
        {
          bulk_g2s(dst, src, bulk_bytes, full_bar + s, policy);
        }
        if (++s == stages)
        {
          s = 0;
          first_pass = false;
          use_parity ^= 1u;
        }
      }
    }
    return;
  }

  if (RIDER && warp == NW + 1)
  {
    // ===== the next view's count stage (smesh_fuse_add_batch) =====
    count_job_warp_call(a.next_ids, a.next_npix, (uint32_t) a.P, a.next_counts, a.next_tag, lane);
    return;
  }

  // ===== consumers: warp cw owns pixels [cw*32, cw*32+32) of every tile, lane = pixel =====
  const int cw = warp - 1;
  const int64_t lane_px = (int64_t) cw * 32 + lane;      // pixel of this lane inside a tile
  const int64_t tile_stride = gridDim.x;
  const uint32_t P32 = (uint32_t) a.P;                   // P < 2^32 - 1 (checked on the host)
  const bool has_weights = a.weights != nullptr;
  const uint32_t run_cap = a.run_cap;                    // power of two: longest run reduced by one reduction

  // Software pipeline of the per-pixel side inputs (all L2 gathers): while tile t is processed, the id / weight of tile
  // t+2 and the pixel count (which needs the id) of tile t+1 are in flight.
  auto load_id = [&](int64_t t) -> uint32_t {
    const int64_t i = t * tile_px + lane_px;
    return (t < a.ntiles && i < a.npix) ? __ldg(a.ids + i) : INVALID_ID;
  };
  auto load_wt = [&](int64_t t) -> float {
    const int64_t i = t * tile_px + lane_px;
    return (has_weights && t < a.ntiles && i < a.npix) ? __ldg(a.weights + i) : 1.0f;
  };
  auto load_n = [&](uint32_t pid) -> uint32_t { return pid < P32 ? __ldg(a.counts + pid) : 1u; };

  int64_t tile = blockIdx.x;
  uint32_t id = load_id(tile), id1 = load_id(tile + tile_stride);
  float wt = load_wt(tile), wt1 = load_wt(tile + tile_stride);
  uint32_t n = load_n(id);

  int s = 0;
  uint32_t parity = 0;
  const float* stage_ptr = stage_base + (size_t) lane_px * C + (PADG > 0 ? (size_t) (lane_px / (PADG > 0 ? PADG : 1)) * 4 : 0);
  for (; tile < a.ntiles; tile += tile_stride)
  {
    const uint32_t id2 = load_id(tile + 2 * tile_stride);
    const float wt2 = load_wt(tile + 2 * tile_stride);
    const uint32_t n1 = load_n(id1);

    const bool valid = id < P32;
    mbar_wait(full_bar + s, parity, a.wait_hint);
    const float* row = stage_ptr + stage_floats * s;

    // ---- gate (Mesh.h:95-98): sequential float sum of the class vector > 0.5 ----
    float v[CH];
    float sum = 0.0f, best = 0.0f;
    int best_c = 0;
    if (C <= CH)
    {
      load_chunk(row, C, al, v); // the whole class vector, zero-padded to CH
      if (KIND == SMESH_KIND_SUMMAX)
      {
        best = v[0];
      }
#pragma unroll
      for (int k = 0; k < CH; k++)
      {
        if (k < C)
        {
          sum = __fadd_rn(sum, v[k]);
          if (KIND == SMESH_KIND_SUMMAX && v[k] > best) // first maximum, strict > (tt/tensor/util/ArgComp.h:3-18)
          {
            best = v[k];
            best_c = k;
          }
        }
      }
    }
    else
    {
      if (KIND == SMESH_KIND_SUMMAX)
      {
        best = row[0];
      }
      for (int c0 = 0; c0 < C; c0 += CH)
      {
        const int nv = min(CH, C - c0);
        load_chunk(row + c0, nv, al, v);
#pragma unroll
        for (int k = 0; k < CH; k++)
        {
          if (k < nv)
          {
            sum = __fadd_rn(sum, v[k]);
            if (KIND == SMESH_KIND_SUMMAX && v[k] > best)
            {
              best = v[k];
              best_c = c0 + k;
            }
          }
        }
      }
    }
    const bool ok = valid && (sum > 0.5f);
    const float w = pixel_weight(a.iew, n & a.count_mask, wt);

    if constexpr (KIND == SMESH_KIND_SUMMAX)
    {
      // one class per pixel (Fusion.cu:51-56): consecutive accepted pixels with the same face AND the same best class
      // fold into one scalar reduction (segmented suffix sum over the run, log2(longest run) shuffle rounds)
      const uint32_t key = ok ? id : INVALID_ID;
      const uint32_t prev = __shfl_up_sync(0xFFFFFFFFu, key, 1);
      const int prev_c = __shfl_up_sync(0xFFFFFFFFu, best_c, 1);
      const bool head = ok && (lane == 0 || prev != key || prev_c != best_c);
      const uint32_t headmask = __ballot_sync(0xFFFFFFFFu, head);
      const uint32_t okmask = __ballot_sync(0xFFFFFFFFu, ok);
      const uint32_t brk = (headmask | ~okmask) & ~((2u << lane) - 1u);
      const int end = ok ? (brk ? (__ffs(brk) - 1) : 32) : lane + 1;
      const int maxlen = (int) __reduce_max_sync(0xFFFFFFFFu, (unsigned) (head ? end - lane : 0));
      float val = ok ? __fmul_rn(best, w) : 0.0f;
      for (int d = 1; d < maxlen; d <<= 1)
      {
        const float t = __shfl_down_sync(0xFFFFFFFFu, val, d);
        if (lane + d < end)
        {
          val = __fadd_rn(val, t);
        }
      }
      if (head)
      {
        red_add_f32(a.acc + (size_t) id * Cpad + best_c, val);
      }
    }
    else
    {
      // ---- runs of equal face id among consecutive accepted pixels of this warp, cut every run_cap lanes ----
      const uint32_t key = ok ? id : INVALID_ID;
      const uint32_t prev = __shfl_up_sync(0xFFFFFFFFu, key, 1);
      const bool head = ok && (lane == 0 || prev != key || (lane & (run_cap - 1)) == 0);
      const uint32_t headmask = __ballot_sync(0xFFFFFFFFu, head);
      const uint32_t okmask = __ballot_sync(0xFFFFFFFFu, ok);
      const uint32_t brk = (headmask | ~okmask) & ~((2u << lane) - 1u);
      const int end = ok ? (brk ? (__ffs(brk) - 1) : 32) : lane + 1; // one past the last pixel of this lane's run
      const int maxlen = (int) __reduce_max_sync(0xFFFFFFFFu, (unsigned) (head ? end - lane : 0));
      float* dst = a.acc + (size_t) (ok ? id : 0) * Cpad;

      for (int c0 = 0; c0 < C; c0 += CH)
      {
        const int nv = min(CH, C - c0);
        if (C > CH)
        {
          load_chunk(row + c0, nv, al, v);
        }
        if (KIND == SMESH_KIND_SUM)
        {
#pragma unroll
          for (int k = 0; k < CH; k++)
          {
            v[k] = __fmul_rn(v[k], w); // weighted::sum (tt/aggregator/MiscOps.h:83-93): acc += probs * w
          }
        }
        else
        {
          // the chunk in the direct form, branch-free; redone element by element if one is outside its domain
          float l[CH];
          bool bad = a.mul_exact != 0 || !(w > 0.0f);
#pragma unroll
          for (int k = 0; k < CH; k++)
          {
            bool bad_k = false;
            l[k] = neg_log_pow_direct(v[k], w, bad_k);
            bad |= bad_k && k < nv;
          }
          if (bad && ok)
          {
#pragma unroll
            for (int k = 0; k < CH; k++)
            {
              l[k] = neg_log_pow_checked(v[k], w, a.mul_exact != 0);
            }
          }
#pragma unroll
          for (int k = 0; k < CH; k++)
          {
            v[k] = (k < nv && ok) ? l[k] : 0.0f;
          }
        }
        // segmented suffix sum over the run (log2(longest run) shuffle rounds, warp-uniform)
        for (int d = 1; d < maxlen; d <<= 1)
        {
          const bool take = lane + d < end;
#pragma unroll
          for (int k = 0; k < CH; k++)
          {
            const float t = __shfl_down_sync(0xFFFFFFFFu, v[k], d);
            if (take)
            {
              v[k] = __fadd_rn(v[k], t);
            }
          }
        }
        // the head of each run owns the run's sum: one 128-bit reduction per 4 classes into the padded row
        if (head)
        {
#pragma unroll
          for (int j = 0; j < CH / 4; j++)
          {
            if (4 * j < nv)
            {
              red_add_v4(dst + c0 + 4 * j, v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            }
          }
        }
      }
    }

    __syncwarp();
    if (lane == 0)
    {
      mbar_arrive(empty_bar + s);
    }
    if (++s == stages)
    {
      s = 0;
      parity ^= 1u;
    }
    id = id1; id1 = id2;
    wt = wt1; wt1 = wt2;
    n = n1;
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// 2'. scatter, narrow class vectors (compile-time C <= 20): TWO adjacent pixels per lane.
//
// Same ring, same arithmetic per pixel; a consumer warp owns 64 consecutive pixels of a stage and lane l holds pixels
// A = 2l and B = 2l+1 in registers (their 2C floats are contiguous and 8-byte aligned: C 64-bit shared loads, conflict
// free). Equal-id neighbours A, B are added in registers first, so the shuffle rounds of the run reduction work on lanes
// that already hold two pixels: about half the shuffles, half the instructions and half the reductions' bookkeeping per
// pixel compared to scatter_kernel. Bookkeeping per lane:
//   S = the part of the lane that belongs to the run entering from the LEFT (A, or A + B when merged)
//   T = B when B starts a new run inside the lane (split lane)
// S-chains are reduced towards their leftmost lane with a segmented suffix scan over lanes; a split lane then adds the
// finished chain of its right neighbour to T. Run heads (S or T) issue the 128-bit reductions.
// ---------------------------------------------------------------------------------------------------------------------

// (A 64-register build for more co-resident rasterizer CTAs was measured in round 2: 94 us per view against 36 - its
// spills sit in the inner loop; profiles/r02e_coresidency_sweep.txt. Removed.)
// Bank conflicts of the lanes' row loads (lane stride 2 C words). Odd C: 64-bit loads, conflict free. C = 2 (mod 4): the 2 C
// floats of a lane are whole 16-byte chunks and 128-bit loads are conflict free. C = 0 (mod 4): 128-bit loads, and every
// group of PADL lanes is shifted by one more 16-byte chunk in shared memory (C = 4, 12, 20: lanes l and l + 4 would meet
// in the same banks, C = 8: l and l + 2, C = 16: all of them); the tile then arrives as one bulk copy per group, issued by
// the lanes of the producer warp in parallel. (Round 1: C = 16 ran at a third of the roofline, 16-way conflicts.)
__host__ __device__ constexpr int pair_pad_lanes(int C)
{
  return (C % 4 != 0) ? 0 : (C % 8 == 4 ? 4 : (C % 16 == 8 ? 2 : 1));
}

// WIDE: the even-C layout (128-bit loads, padded groups); false = 64-bit loads of the dense tile for every C
template <int KIND, int CT, bool WIDE = true>
__global__ void __maxnreg__(88) scatter_pair_kernel(ScatterArgs a) // (launched with at most 320 threads)
{
  static_assert(CT >= 1 && CT <= CH, "pair kernel holds 2 x Cpad values in registers");
  constexpr int C = CT;
  constexpr int Cpad = (CT + 3) & ~3;
  constexpr int NCHUNK = Cpad / 4;
  constexpr int PADL = WIDE ? pair_pad_lanes(CT) : 0;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const bool has_count = a.next_ids != nullptr;
  const int NW = (int) (blockDim.x >> 5) - 1 - (has_count ? 1 : 0); // consumer warps (warp 0 produces, the last may count)
  const int tile_px = NW * 64;
  const size_t stage_floats = (size_t) tile_px * C + (PADL > 0 ? (size_t) (NW * 32 / (PADL > 0 ? PADL : 1)) * 4 : 0);
  const int stages = a.stages;

  float* stage_base = reinterpret_cast<float*>(smem_raw);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(stage_base + stage_floats * stages);
  uint64_t* empty_bar = full_bar + stages;
  // per consumer warp 64 row slots of Cpad floats + their face ids: the run sums staged for the coalesced flush
  float* flush_base = reinterpret_cast<float*>(empty_bar + stages);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0)
  {
    for (int s = 0; s < stages; s++)
    {
      mbar_init(full_bar + s, 1);
      mbar_init(empty_bar + s, (uint32_t) NW);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (warp == 0 && PADL > 0)
  {
    // ===== producer, padded layout: one bulk copy per group of PADL lanes (2 PADL pixels), issued by the lanes in parallel =====
    const uint64_t policy = l2_evict_first_policy();
    constexpr int GP = 2 * (PADL > 0 ? PADL : 1); // pixels per group
    int s = 0;
    uint32_t use_parity = 1;
    bool first_pass = true;
    for (int64_t tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x)
    {
      if (!first_pass)
      {
        if (lane == 0)
        {
          mbar_wait(empty_bar + s, use_parity, a.wait_hint);
        }
        __syncwarp();
      }
      const int64_t px0 = tile * tile_px;
      const int px_n = (int) min((int64_t) tile_px, a.npix - px0);
      float* dst = stage_base + stage_floats * s;
      const float* src = a.probs + (size_t) px0 * C;
      if (lane == 0)
      {
        mbar_arrive_expect_tx(full_bar + s, (uint32_t) ((size_t) px_n * C * 4)); // C % 4 == 0: whole 16-byte chunks
      }
      __syncwarp();
      const int ngroups = (px_n + GP - 1) / GP;
      for (int g = lane; g < ngroups; g += 32)
      {
        const int gp = min(GP, px_n - g * GP);
        bulk_g2s(dst + (size_t) g * (GP * C + 4), src + (size_t) g * GP * C, (uint32_t) (gp * C * 4), full_bar + s, policy);
      }
      if (++s == stages)
      {
        s = 0;
        first_pass = false;
        use_parity ^= 1u;
      }
    }
    return;
  }
  if (warp == 0)
  {
    if (lane == 0)
    {
      const uint64_t policy = l2_evict_first_policy();
      int s = 0;
      uint32_t use_parity = 1;
      bool first_pass = true;
      for (int64_t tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x)
      {
        if (!first_pass)
        {
          mbar_wait(empty_bar + s, use_parity, a.wait_hint);
        }
        const int64_t px0 = tile * tile_px;
        const int64_t px_n = min((int64_t) tile_px, a.npix - px0);
        const size_t nfloats = (size_t) px_n * C;
        const uint32_t bulk_bytes = (uint32_t) ((nfloats * 4) & ~(size_t) 15);
        float* dst = stage_base + stage_floats * s;
        const float* src = a.probs + (size_t) px0 * C;
        for (size_t k = bulk_bytes / 4; k < nfloats; k++)
        {
          dst[k] = src[k];
        }
        mbar_arrive_expect_tx(full_bar + s, bulk_bytes);
        if (bulk_bytes > 0)
        {
          bulk_g2s(dst, src, bulk_bytes, full_bar + s, policy);
        }
        if (++s == stages)
        {
          s = 0;
          first_pass = false;
          use_parity ^= 1u;
        }
      }
    }
    return;
  }

  if (warp == NW + 1)
  {
    // ===== the next view's count stage (smesh_fuse_add_batch): hidden under this view's scatter =====
    count_job_warp(a.next_ids, a.next_npix, (uint32_t) a.P, a.next_counts, a.next_tag, lane);
    return;
  }

  const int cw = warp - 1;
  const int64_t lane_px = (int64_t) cw * 64 + 2 * lane; // first of this lane's two pixels inside a tile
  const int64_t tile_stride = gridDim.x;
  const uint32_t P32 = (uint32_t) a.P;
  const bool has_weights = a.weights != nullptr;

  // side inputs (L2 gathers) are fetched two tiles ahead (ids, weights) / one tile ahead (counts, which need the ids)
  auto load_ids = [&](int64_t t) -> uint2 {
    const int64_t i = t * tile_px + lane_px;
    uint2 r = make_uint2(INVALID_ID, INVALID_ID);
    if (i + 1 < a.npix)
    {
      r = __ldg(reinterpret_cast<const uint2*>(a.ids + i)); // i is even and the id image is 8-byte aligned (checked on host)
    }
    else if (i < a.npix)
    {
      r.x = __ldg(a.ids + i);
    }
    return r;
  };
  auto load_wts = [&](int64_t t) -> float2 {
    const int64_t i = t * tile_px + lane_px;
    float2 r = make_float2(1.0f, 1.0f);
    if (has_weights)
    {
      if (i < a.npix) r.x = __ldg(a.weights + i);
      if (i + 1 < a.npix) r.y = __ldg(a.weights + i + 1);
    }
    return r;
  };
  auto load_n = [&](uint32_t pid) -> uint32_t { return pid < P32 ? __ldg(a.counts + pid) : 1u; };

  int64_t tile = blockIdx.x;
  uint2 id = load_ids(tile), id1 = load_ids(tile + tile_stride);
  float2 wt = load_wts(tile), wt1 = load_wts(tile + tile_stride);
  uint2 n = make_uint2(load_n(id.x), load_n(id.y));

  int s = 0;
  uint32_t parity = 0;
  const float* stage_ptr = stage_base + (size_t) lane_px * C + (PADL > 0 ? (size_t) ((cw * 32 + lane) / (PADL > 0 ? PADL : 1)) * 4 : 0);
  for (; tile < a.ntiles; tile += tile_stride)
  {
    const uint2 id2 = load_ids(tile + 2 * tile_stride);
    const float2 wt2 = load_wts(tile + 2 * tile_stride);
    const uint2 n1 = make_uint2(load_n(id1.x), load_n(id1.y));

    mbar_wait(full_bar + s, parity, a.wait_hint);
    // ---- both pixels' class vectors: 2C contiguous floats, C 64-bit loads (even C: C / 2 128-bit loads) ----
    float ab[2 * C];
    if constexpr (C % 2 == 0 && WIDE)
    {
      const float4* row4 = reinterpret_cast<const float4*>(stage_ptr + stage_floats * s);
#pragma unroll
      for (int k = 0; k < C / 2; k++)
      {
        const float4 t = row4[k];
        ab[4 * k] = t.x;
        ab[4 * k + 1] = t.y;
        ab[4 * k + 2] = t.z;
        ab[4 * k + 3] = t.w;
      }
    }
    else
    {
      const float2* row2 = reinterpret_cast<const float2*>(stage_ptr + stage_floats * s);
#pragma unroll
      for (int k = 0; k < C; k++)
      {
        const float2 t = row2[k];
        ab[2 * k] = t.x;
        ab[2 * k + 1] = t.y;
      }
    }
    float A[Cpad], B[Cpad];
#pragma unroll
    for (int k = 0; k < Cpad; k++)
    {
      A[k] = k < C ? ab[k] : 0.0f;
      B[k] = k < C ? ab[C + k] : 0.0f;
    }
    // ---- gate (Mesh.h:95-98): sequential float sum of each class vector > 0.5 ----
    float sumA = 0.0f, sumB = 0.0f;
#pragma unroll
    for (int k = 0; k < C; k++)
    {
      sumA = __fadd_rn(sumA, A[k]);
      sumB = __fadd_rn(sumB, B[k]);
    }
    const bool okA = id.x < P32 && sumA > 0.5f;
    const bool okB = id.y < P32 && sumB > 0.5f;
    const float wA = pixel_weight(a.iew, n.x & a.count_mask, wt.x);
    const float wB = pixel_weight(a.iew, n.y & a.count_mask, wt.y);
    if (KIND == SMESH_KIND_SUM)
    {
#pragma unroll
      for (int k = 0; k < Cpad; k++)
      {
        A[k] = __fmul_rn(A[k], wA); // weighted::sum (tt/aggregator/MiscOps.h:83-93): acc += probs * w
        B[k] = __fmul_rn(B[k], wB);
      }
    }
    else
    {
      // both vectors in the direct form, branch-free; a vector with an element outside its domain is redone element by
      // element from the stage (still owned by this warp until the end of the iteration)
      bool badA = a.mul_exact != 0 || !(wA > 0.0f), badB = a.mul_exact != 0 || !(wB > 0.0f);
#pragma unroll
      for (int k = 0; k < C; k++)
      {
        A[k] = neg_log_pow_direct(A[k], wA, badA);
        B[k] = neg_log_pow_direct(B[k], wB, badB);
      }
      badA = badA && okA;
      badB = badB && okB;
      if (badA || badB)
      {
        const float* rowf = stage_ptr + stage_floats * s;
#pragma unroll
        for (int k = 0; k < C; k++)
        {
          if (badA) A[k] = neg_log_pow_checked(rowf[k], wA, a.mul_exact != 0);
          if (badB) B[k] = neg_log_pow_checked(rowf[C + k], wB, a.mul_exact != 0);
        }
      }
#pragma unroll
      for (int k = 0; k < C; k++)
      {
        A[k] = okA ? A[k] : 0.0f;
        B[k] = okB ? B[k] : 0.0f;
      }
    }

    // ---- run structure ----
    const uint32_t keyA = okA ? id.x : INVALID_ID, keyB = okB ? id.y : INVALID_ID;
    const bool merged = okA && okB && id.x == id.y;
    const bool hasT = okB && !merged;
    if (merged)
    {
#pragma unroll
      for (int k = 0; k < Cpad; k++)
      {
        A[k] = __fadd_rn(A[k], B[k]); // S = A + B
      }
    }
    const uint32_t keyR = merged ? id.x : keyB;                       // id of the run leaving the lane on the right
    const uint32_t keyA_next = __shfl_down_sync(0xFFFFFFFFu, keyA, 1);  // lane + 1's left pixel
    const uint32_t keyR_prev = __shfl_up_sync(0xFFFFFFFFu, keyR, 1);    // lane - 1's right pixel
    const bool joins_next = lane < 31 && keyR != INVALID_ID && keyA_next == keyR;
    const bool contS = merged && joins_next;    // the S chain continues into lane + 1
    const bool contT = hasT && joins_next;      // T (= B) is the head of a run that continues into lane + 1
    const bool absorbed = lane > 0 && okA && keyR_prev == keyA; // A belongs to a run whose head is further left
    const bool headS = okA && !absorbed;
    const uint32_t contmask = __ballot_sync(0xFFFFFFFFu, contS);
    const uint32_t stop = ~contmask & ~((1u << lane) - 1u);           // first lane >= this one whose S chain stops
    const int end = __ffs(stop);                                        // = that lane + 1 (lane 31 never continues)
    const int maxlen = (int) __reduce_max_sync(0xFFFFFFFFu, (unsigned) (end - lane));

    // segmented suffix sum of S over the chain (log2(longest chain) shuffle rounds, warp-uniform)
    for (int d = 1; d < maxlen; d <<= 1)
    {
      const bool take = lane + d < end;
#pragma unroll
      for (int k = 0; k < Cpad; k++)
      {
        const float t = __shfl_down_sync(0xFFFFFFFFu, A[k], d);
        if (take)
        {
          A[k] = __fadd_rn(A[k], t);
        }
      }
    }
    // a split lane's B heads the run whose remainder is the finished chain of lane + 1
    if (__any_sync(0xFFFFFFFFu, contT))
    {
#pragma unroll
      for (int k = 0; k < Cpad; k++)
      {
        const float t = __shfl_down_sync(0xFFFFFFFFu, A[k], 1);
        if (contT)
        {
          B[k] = __fadd_rn(B[k], t);
        }
      }
    }
    // run heads own the sums and add them to the padded accumulator row (Cpad floats, 16-byte aligned)
    {
      // Coalesced flush: the run sums are compacted into shared memory rows, then every group of 5 adjacent lanes adds
      // one 80-byte row with a single red.global.add.v4.f32 instruction (6 rows per warp instruction): the L2 receives
      // whole-sector requests (3 per row instead of 5 half-sector ones) and no lane idles in a per-row loop.
      float* rows = flush_base + (size_t) cw * (64 * Cpad + 64);
      uint32_t* row_id = reinterpret_cast<uint32_t*>(rows + 64 * Cpad);
      const bool doS = headS, doT = hasT;
      const uint32_t smask = __ballot_sync(0xFFFFFFFFu, doS), tmask = __ballot_sync(0xFFFFFFFFu, doT);
      const uint32_t below = (1u << lane) - 1u;
      const int nS = __popc(smask), nrows = nS + __popc(tmask);
      if (doS)
      {
        const int r = __popc(smask & below);
        float* dst = rows + r * Cpad;
#pragma unroll
        for (int j = 0; j < NCHUNK; j++)
        {
          *reinterpret_cast<float4*>(dst + 4 * j) = make_float4(A[4 * j], A[4 * j + 1], A[4 * j + 2], A[4 * j + 3]);
        }
        row_id[r] = id.x;
      }
      if (doT)
      {
        const int r = nS + __popc(tmask & below);
        float* dst = rows + r * Cpad;
#pragma unroll
        for (int j = 0; j < NCHUNK; j++)
        {
          *reinterpret_cast<float4*>(dst + 4 * j) = make_float4(B[4 * j], B[4 * j + 1], B[4 * j + 2], B[4 * j + 3]);
        }
        row_id[r] = id.y;
      }
      __syncwarp();
      constexpr int GROUPS = 32 / NCHUNK;          // rows per warp instruction
      const int sub = lane / NCHUNK, j = lane - sub * NCHUNK;
      if (sub < GROUPS)
      {
        for (int r = sub; r < nrows; r += GROUPS)
        {
          const float4 v = *reinterpret_cast<const float4*>(rows + r * Cpad + 4 * j);
          red_add_v4(a.acc + (size_t) row_id[r] * Cpad + 4 * j, v.x, v.y, v.z, v.w);
        }
      }
      __syncwarp();
    }

    __syncwarp();
    if (lane == 0)
    {
      mbar_arrive(empty_bar + s);
    }
    if (++s == stages)
    {
      s = 0;
      parity ^= 1u;
    }
    id = id1; id1 = id2;
    wt = wt1; wt1 = wt2;
    n = n1;
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// 2c. scatter, WIDE class vectors (C >= 32): lanes across the classes of one pixel.
//
// A warp takes 32 consecutive pixels: lane = pixel for the side inputs (id, per-face count, weight); then the lanes turn
// across the class vector: a sub-group of LG = 8 / 16 / 32 lanes reads a pixel's row straight from global memory with
// 128- or 64-bit loads (consecutive lanes, consecutive classes: fully coalesced, streaming, no shared memory), weights it
// and keeps the sum of the current run of equal face ids in registers, walking its share of the 32 pixels in order; a run
// ends with one vector reduction per lane into the accumulator row (again consecutive lanes, consecutive addresses).
// No shuffle scans, no staging, no divergence inside a sub-group.
// The gate (Mesh.h:98) needs the SEQUENTIAL float sum of the row. The lanes form the sum in tree order; for
// non-negative values the two orders differ by at most 2 C 2^-24 of the sum, so unless the tree sum is within 1e-3
// (relative) of 0.5 - or a value is negative or NaN - the comparison `> 0.5` has the same outcome; otherwise one lane
// recomputes the sum in the reference's order.
// ---------------------------------------------------------------------------------------------------------------------

template <int VW>
__device__ __forceinline__ void row_load(const float* p, float (&v)[VW])
{
  if (VW == 4)
  {
    const float4 t = __ldcs(reinterpret_cast<const float4*>(p));
    v[0] = t.x; v[1 % VW] = t.y; v[2 % VW] = t.z; v[3 % VW] = t.w;
  }
  else if (VW == 2)
  {
    const float2 t = __ldcs(reinterpret_cast<const float2*>(p));
    v[0] = t.x; v[1 % VW] = t.y;
  }
  else
  {
    v[0] = __ldcs(p);
  }
}

template <int VW>
__device__ __forceinline__ void row_red(float* p, const float (&v)[VW])
{
  if (VW == 4)
  {
    red_add_v4(p, v[0], v[1 % VW], v[2 % VW], v[3 % VW]);
  }
  else if (VW == 2)
  {
    asm volatile("red.relaxed.gpu.global.add.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(v[0]), "f"(v[1 % VW]) : "memory");
  }
  else
  {
    red_add_f32(p, v[0]);
  }
}

// VW = floats per lane and chunk (C % VW == 0), LG = lanes of a sub-group, NP = chunks per lane (ceil(C / VW / LG))
template <int KIND, int VW, int LG, int NP>
__global__ void __launch_bounds__(256) scatter_rows_kernel(ScatterArgs a)
{
  constexpr int U = NP <= 2 ? 4 : 2; // pixels whose rows are loaded before the first is reduced
  constexpr int G = 32 / LG;         // sub-groups
  constexpr int Rg = 32 / G;         // pixels of a sub-group per warp block
  const int lane = threadIdx.x & 31;
  const int g = lane / LG, j = lane % LG;
  const int K = a.C / VW;                        // chunks per row
  const uint32_t P32 = (uint32_t) a.P;
  const int64_t warp_global = ((int64_t) blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t) gridDim.x * blockDim.x) >> 5;

  auto load_side = [&](int64_t base, uint32_t& id, float& wt) {
    const int64_t i = base + lane;
    id = INVALID_ID;
    wt = 1.0f;
    if (i < a.npix)
    {
      id = __ldg(a.ids + i);
      if (a.weights != nullptr)
      {
        wt = __ldg(a.weights + i);
      }
    }
    if (!(id < P32))
    {
      id = INVALID_ID;
    }
  };

  int64_t base = warp_global * 32;
  uint32_t id_next;
  float wt_next;
  load_side(base, id_next, wt_next);
  for (; base < a.npix; base += nwarps * 32)
  {
    const uint32_t id = id_next;
    const float wt = wt_next;
    const uint32_t n = id != INVALID_ID ? __ldg(a.counts + id) : 1u;
    load_side(base + nwarps * 32, id_next, wt_next); // the next block's ids are in flight while this one is reduced
    const float w = pixel_weight(a.iew, n & a.count_mask, wt);

    uint32_t cur = INVALID_ID;
    float acc[NP][VW];
#pragma unroll
    for (int p = 0; p < NP; p++)
    {
#pragma unroll
      for (int k = 0; k < VW; k++)
      {
        acc[p][k] = 0.0f;
      }
    }
    auto flush = [&]() {
      if (cur != INVALID_ID)
      {
        float* dst = a.acc + (size_t) cur * a.Cpad;
#pragma unroll
        for (int p = 0; p < NP; p++)
        {
          const int c = j + p * LG;
          if (c < K)
          {
            row_red<VW>(dst + c * VW, acc[p]);
          }
#pragma unroll
          for (int k = 0; k < VW; k++)
          {
            acc[p][k] = 0.0f;
          }
        }
      }
    };

    // this lane's first chunk of the sub-group's first pixel (Rg and U divide evenly: no pixel is out of range)
    const float* lane_row = a.probs + (size_t) base * a.C + (g * Rg) * a.C + j * VW;
    static_assert(Rg % U == 0, "sub-group pixels come in whole batches");
    for (int it = 0; it < Rg; it += U)
    {
      uint32_t idu[U];
      float wu[U];
      float v[U][NP][VW];
#pragma unroll
      for (int u = 0; u < U; u++)
      {
        const int q = g * Rg + it + u;           // pixel of the warp block this sub-group looks at
        idu[u] = __shfl_sync(0xFFFFFFFFu, id, q);
        wu[u] = __shfl_sync(0xFFFFFFFFu, w, q);
        const float* row = lane_row + (it + u) * a.C;
#pragma unroll
        for (int p = 0; p < NP; p++)
        {
#pragma unroll
          for (int k = 0; k < VW; k++)
          {
            v[u][p][k] = 0.0f;
          }
          if (idu[u] != INVALID_ID && j + p * LG < K)
          {
            row_load<VW>(row + p * (LG * VW), v[u][p]);
          }
        }
      }
#pragma unroll
      for (int u = 0; u < U; u++)
      {
        // ---- gate (Mesh.h:95-98) ----
        float s = 0.0f;
        uint32_t signs = 0u; // a negative value: the tree sum says nothing about the sequential one (a NaN shows in s)
#pragma unroll
        for (int p = 0; p < NP; p++)
        {
#pragma unroll
          for (int k = 0; k < VW; k++)
          {
            s += v[u][p][k];
            signs |= __float_as_uint(v[u][p][k]);
          }
        }
#pragma unroll
        for (int o = LG >> 1; o > 0; o >>= 1)
        {
          s += __shfl_xor_sync(0xFFFFFFFFu, s, o);
        }
        const uint32_t oddmask = __ballot_sync(0xFFFFFFFFu, (signs & 0x80000000u) != 0u);
        const uint32_t gmask = (LG == 32 ? 0xFFFFFFFFu : ((1u << (LG % 32)) - 1u)) << (g * LG);
        const bool unsure = (oddmask & gmask) != 0u || !(fabsf(s - 0.5f) > 1e-3f * fmaxf(s, 0.5f));
        if (idu[u] != INVALID_ID && unsure)
        {
          // the reference's order, by the first lane of the sub-group (rare: a sum at the threshold, or odd values)
          float seq = 0.0f;
          if (j == 0)
          {
            const float* row = a.probs + (size_t) (base + g * Rg + it + u) * a.C;
#pragma unroll 1
            for (int c = 0; c < a.C; c++)
            {
              seq = __fadd_rn(seq, row[c]);
            }
          }
          s = seq;
        }
        s = __shfl_sync(0xFFFFFFFFu, s, g * LG); // (uniform inside the sub-group either way)
        const bool ok = idu[u] != INVALID_ID && s > 0.5f;
        if (ok)
        {
          if (idu[u] != cur)
          {
            flush();
            cur = idu[u];
          }
          if (KIND == SMESH_KIND_SUM)
          {
#pragma unroll
            for (int p = 0; p < NP; p++)
            {
#pragma unroll
              for (int k = 0; k < VW; k++)
              {
                acc[p][k] = __fadd_rn(acc[p][k], __fmul_rn(v[u][p][k], wu[u])); // acc += probs * w (MiscOps.h:83-93)
              }
            }
          }
          else
          {
            // this lane's elements in the direct form, branch-free; redone one by one if one is outside its domain
            float l[NP][VW];
            bool bad = a.mul_exact != 0 || !(wu[u] > 0.0f);
#pragma unroll
            for (int p = 0; p < NP; p++)
            {
#pragma unroll
              for (int k = 0; k < VW; k++)
              {
                bool bad_k = false;
                l[p][k] = neg_log_pow_direct(v[u][p][k], wu[u], bad_k);
                bad |= bad_k && j + p * LG < K;
              }
            }
            if (bad)
            {
#pragma unroll
              for (int p = 0; p < NP; p++)
              {
#pragma unroll
                for (int k = 0; k < VW; k++)
                {
                  l[p][k] = neg_log_pow_checked(v[u][p][k], wu[u], a.mul_exact != 0);
                }
              }
            }
#pragma unroll
            for (int p = 0; p < NP; p++)
            {
#pragma unroll
              for (int k = 0; k < VW; k++)
              {
                if (j + p * LG < K)
                {
                  acc[p][k] = __fadd_rn(acc[p][k], l[p][k]);
                }
              }
            }
          }
        }
      }
    }
    flush();
  }
}

template <int KIND, int VW>
static int launch_scatter_rows(const ScatterArgs& args, cudaStream_t stream)
{
  const int K = args.C / VW;
  const int lg = K <= 8 ? 8 : (K <= 16 ? 16 : 32);
  const int np = (K + lg - 1) / lg;
  int64_t blocks = (args.npix + 255) / 256;
  const int64_t cap = (int64_t) num_sms() * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) return SMESH_OK;
  const unsigned nb = (unsigned) blocks;
  if (lg == 8) scatter_rows_kernel<KIND, VW, 8, 1><<<nb, 256, 0, stream>>>(args);
  else if (lg == 16) scatter_rows_kernel<KIND, VW, 16, 1><<<nb, 256, 0, stream>>>(args);
  else if (np == 1) scatter_rows_kernel<KIND, VW, 32, 1><<<nb, 256, 0, stream>>>(args);
  else if (np == 2) scatter_rows_kernel<KIND, VW, 32, 2><<<nb, 256, 0, stream>>>(args);
  else if (np == 3) scatter_rows_kernel<KIND, VW, 32, 3><<<nb, 256, 0, stream>>>(args);
  else if (np == 4) scatter_rows_kernel<KIND, VW, 32, 4><<<nb, 256, 0, stream>>>(args);
  else
  {
    set_error("scatter_rows_kernel: class vector too wide (C=%d)", args.C);
    return SMESH_ERR_UNSUPPORTED;
  }
  SMESH_LAUNCH_CHECK("scatter_rows_kernel");
  return SMESH_OK;
}

// C >= 32 with 16 or more chunks per row and at most 4 per lane, rows aligned for the vector width (C % 4 == 0 with a 16-byte aligned image,
// C % 2 == 0 with an 8-byte aligned one, any alignment for odd C)
static bool rows_kernel_takes(const ScatterArgs& args, int& vw)
{
  static const bool off = getenv("SMESH_NO_ROWS") != nullptr;
  if (off || args.C < 32)
  {
    return false;
  }
  const uintptr_t addr = reinterpret_cast<uintptr_t>(args.probs);
  vw = (args.C % 4 == 0 && (addr & 15) == 0) ? 4 : ((args.C % 2 == 0 && (addr & 7) == 0) ? 2 : 1);
  // at least 16 chunks per row (full sub-groups; measured at C = 40, 10 chunks: 29 us against 25 us for the ring kernel)
  return args.C / vw >= 16 && (args.C / vw + 31) / 32 <= 4;
}

// Fallback for shapes the ring cannot take (class vector too wide for shared memory, misaligned probability image):
// one thread per pixel straight from global memory. Same arithmetic, no staging.
template <int KIND>
__global__ void __launch_bounds__(256) scatter_direct_kernel(ScatterArgs a)
{
  const int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.npix)
  {
    return;
  }
  const uint32_t id = a.ids[i];
  if (!((int64_t) id < a.P))
  {
    return;
  }
  const float* row = a.probs + (size_t) i * a.C;
  float sum = 0.0f, best = row[0];
  int best_c = 0;
  for (int c = 0; c < a.C; c++)
  {
    const float p = row[c];
    sum = __fadd_rn(sum, p);
    if (p > best)
    {
      best = p;
      best_c = c;
    }
  }
  if (!(sum > 0.5f))
  {
    return;
  }
  const float w = pixel_weight(a.iew, a.counts[id] & a.count_mask, a.weights ? a.weights[i] : 1.0f);
  float* dst = a.acc + (size_t) id * a.Cpad;
  if (KIND == SMESH_KIND_SUMMAX)
  {
    red_add_f32(dst + best_c, __fmul_rn(best, w));
    return;
  }
  for (int c = 0; c < a.C; c++)
  {
    red_add_f32(dst + c, KIND == SMESH_KIND_SUM ? __fmul_rn(row[c], w) : neg_log_pow(row[c], w, a.mul_exact != 0));
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// get(): per-face class distribution
// ---------------------------------------------------------------------------------------------------------------------

// One accumulator row -> its distribution, in place (row[0 .. C) with element stride 1): mul: exp(-(l - min l)), then the
// reference's sequential L1 norm, v * (1 / norm), NaN/Inf -> 0 (Fusion.cu:66-92, Fusion.h:79-104).
template <int KIND>
__device__ __forceinline__ void get_row(float* row, int C)
{
  float best = CUDART_INF_F;
  bool have = false;
  if (KIND == SMESH_KIND_MUL)
  {
    // max_el over LogProb = the smallest -log that is not "zero" (isinf), tt/numeric/LogProb.h
    for (int c = 0; c < C; c++)
    {
      const float v = row[c];
      if (!isinf(v) && (!have || v < best))
      {
        best = v;
        have = true;
      }
    }
  }
  // l1 norm: sequential sum of |v| from 0 (tt/tensor/linear_algebra/MiscOps.h:121-128)
  float norm = 0.0f;
  for (int c = 0; c < C; c++)
  {
    float v = row[c];
    if (KIND == SMESH_KIND_MUL)
    {
      v = (!have || isinf(v)) ? 0.0f : expf(-__fsub_rn(v, best));
      row[c] = v;
    }
    norm = __fadd_rn(norm, fabsf(v));
  }
  const float inv = __fdiv_rn(1.0f, norm);
  for (int c = 0; c < C; c++)
  {
    const float x = __fmul_rn(row[c], inv);
    row[c] = (isnan(x) || isinf(x)) ? 0.0f : x; // Fusion.h:79-95
  }
}

// get(), staged (mul with C > 24, any kind with 200 < C <~ 1500, unaligned outputs; the streaming kernel below takes the
// rest): HBM-bound (read P x Cpad, write P x C floats). A CTA stages `rows` accumulator rows in shared memory with
// 128-bit coalesced loads (row stride Cpad + 1 words: a thread per row then walks its row without bank conflicts), one
// thread per row turns it into the distribution in place in the reference's sequential order, and the block of rows
// leaves as one contiguous run of rows * C floats with 128-bit coalesced stores.
template <int KIND>
__global__ void __launch_bounds__(256) get_kernel(const float* __restrict__ acc, int64_t P, int C, int Cpad, int rows,
                                                  float* __restrict__ out)
{
  extern __shared__ __align__(16) float get_smem[];
  const int stride = Cpad + 1;
  // e / C and i / chunks for e, i < 2^16 (a block of rows has at most 200 KB / 4 elements) as one multiply-high each:
  // an integer division per element made this copy kernel issue-bound
  const uint32_t magic_c = 0xFFFFFFFFu / (uint32_t) C + 1u, magic_chunks = 0xFFFFFFFFu / (uint32_t) (Cpad >> 2) + 1u;
  const int64_t nblocks = (P + rows - 1) / rows;
  for (int64_t blk = blockIdx.x; blk < nblocks; blk += gridDim.x)
  {
    const int64_t r0 = blk * rows;
    const int n = (int) min((int64_t) rows, P - r0);
    const int chunks = Cpad >> 2;                       // float4 per row
    const float4* src = reinterpret_cast<const float4*>(acc + (size_t) r0 * Cpad);
    for (int i = threadIdx.x; i < n * chunks; i += blockDim.x)
    {
      const float4 v = __ldcs(src + i);
      const int r = chunks == 1 ? i : (int) __umulhi((uint32_t) i, magic_chunks), j = i - r * chunks;
      float* d = get_smem + r * stride + 4 * j;
      d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
    }
    __syncthreads();
    for (int r = threadIdx.x; r < n; r += blockDim.x)
    {
      get_row<KIND>(get_smem + r * stride, C);
    }
    __syncthreads();
    float* dst = out + (size_t) r0 * C;
    const int total = n * C;
    if (((reinterpret_cast<uintptr_t>(dst) & 15) == 0))
    {
      const int total4 = total >> 2;
      for (int i = threadIdx.x; i < total4; i += blockDim.x)
      {
        float v[4];
#pragma unroll
        for (int k = 0; k < 4; k++)
        {
          const int e = 4 * i + k;
          const int r = C == 1 ? e : (int) __umulhi((uint32_t) e, magic_c);
          v[k] = get_smem[r * stride + (e - r * C)];
        }
        __stcs(reinterpret_cast<float4*>(dst) + i, make_float4(v[0], v[1], v[2], v[3]));
      }
      for (int e = (total4 << 2) + threadIdx.x; e < total; e += blockDim.x)
      {
        const int r = C == 1 ? e : (int) __umulhi((uint32_t) e, magic_c);
        dst[e] = get_smem[r * stride + (e - r * C)];
      }
    }
    else
    {
      for (int e = threadIdx.x; e < total; e += blockDim.x)
      {
        const int r = C == 1 ? e : (int) __umulhi((uint32_t) e, magic_c);
        dst[e] = get_smem[r * stride + (e - r * C)];
      }
    }
    __syncthreads();
  }
}

// get() as a stream of bulk copies (C <= 200). Every WARP runs its own two-stage pipeline over blocks of 32 rows: one
// bulk copy brings the block's 32 x Cpad floats into shared memory (mbarrier complete_tx), lane r turns row r into its
// distribution - read with 128-bit shared loads (conflict-free when Cpad / 4 is odd, as for C = 19), in the reference's
// sequential order - and writes it into a dense 32 x C staging block, which leaves as one shared -> global bulk copy
// (bulk group; its buffer is reused once `wait_group.read` says it has been read). No __syncthreads, no address
// arithmetic per element, the loads of the next two blocks always in flight.
template <int VW>
__device__ __forceinline__ void put_row(float* dst, const float (&v)[4], int c, int C)
{
  if (VW == 4)
  {
    *reinterpret_cast<float4*>(dst + c) = make_float4(v[0], v[1], v[2], v[3]); // (C % 4 == 0: whole chunks only)
  }
  else if (VW == 2)
  {
    *reinterpret_cast<float2*>(dst + c) = make_float2(v[0], v[1]);
    if (c + 2 < C) *reinterpret_cast<float2*>(dst + c + 2) = make_float2(v[2], v[3]);
  }
  else
  {
#pragma unroll
    for (int k = 0; k < 4; k++)
    {
      if (c + k < C) dst[c + k] = v[k];
    }
  }
}

template <int KIND, int VW>
__global__ void __launch_bounds__(128) get_stream_kernel(const float* __restrict__ acc, int64_t P, int C, int Cpad,
                                                         float* __restrict__ out)
{
  extern __shared__ __align__(128) uint8_t get_stream_smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int in_floats = 32 * Cpad, out_floats = 32 * C;            // (both multiples of 32 floats = 128 bytes)
  float* base = reinterpret_cast<float*>(get_stream_smem) + (size_t) warp * 2 * (in_floats + out_floats);
  float* in_buf[2] = {base, base + in_floats};
  float* out_buf[2] = {base + 2 * in_floats, base + 2 * in_floats + out_floats};
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<float*>(get_stream_smem) + (size_t) nwarps * 2 * (in_floats + out_floats)) + 2 * warp;

  const int64_t nblocks = (P + 31) / 32;
  const int64_t stride = (int64_t) gridDim.x * nwarps;
  int64_t blk = (int64_t) blockIdx.x * nwarps + warp;
  const uint64_t policy = l2_evict_first_policy();
  auto fetch = [&](int64_t b, int s) {
    const int n = (int) min((int64_t) 32, P - b * 32);
    const uint32_t bytes = (uint32_t) n * (uint32_t) Cpad * 4u;
    mbar_arrive_expect_tx(bars + s, bytes);
    bulk_g2s(in_buf[s], acc + (size_t) b * 32 * Cpad, bytes, bars + s, policy);
  };
  if (lane == 0)
  {
    mbar_init(bars, 1);
    mbar_init(bars + 1, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    if (blk < nblocks) fetch(blk, 0);
    if (blk + stride < nblocks) fetch(blk + stride, 1);
  }
  __syncwarp();

  const int chunks = Cpad >> 2;
  uint32_t parity = 0;
  int s = 0;
  for (; blk < nblocks; blk += stride)
  {
    const int n = (int) min((int64_t) 32, P - blk * 32);
    mbar_wait(bars + s, parity);
    if (lane == 0)
    {
      bulk_wait_group_read<1>(); // the copy that left out_buf[s] two blocks ago has read it
    }
    __syncwarp();
    if (lane < n)
    {
      float4* row = reinterpret_cast<float4*>(in_buf[s] + lane * Cpad);
      float best = CUDART_INF_F;
      bool have = false;
      if (KIND == SMESH_KIND_MUL)
      {
        // max_el over LogProb = the smallest -log that is not "zero" (isinf), tt/numeric/LogProb.h
        for (int j = 0; j < chunks; j++)
        {
          const float4 t = row[j];
          const float v[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
          for (int k = 0; k < 4; k++)
          {
            if (4 * j + k < C && !isinf(v[k]) && (!have || v[k] < best))
            {
              best = v[k];
              have = true;
            }
          }
        }
      }
      // l1 norm: sequential sum of |v| from 0 (tt/tensor/linear_algebra/MiscOps.h:121-128)
      float norm = 0.0f;
      for (int j = 0; j < chunks; j++)
      {
        const float4 t = row[j];
        float v[4] = {t.x, t.y, t.z, t.w};
        if (KIND == SMESH_KIND_MUL)
        {
#pragma unroll
          for (int k = 0; k < 4; k++)
          {
            v[k] = (!have || isinf(v[k])) ? 0.0f : expf(-__fsub_rn(v[k], best));
          }
          row[j] = make_float4(v[0], v[1], v[2], v[3]);
        }
#pragma unroll
        for (int k = 0; k < 4; k++)
        {
          if (4 * j + k < C) norm = __fadd_rn(norm, fabsf(v[k]));
        }
      }
      const float inv = __fdiv_rn(1.0f, norm);
      float* dst = out_buf[s] + lane * C;
      for (int j = 0; j < chunks; j++)
      {
        const float4 t = row[j];
        float v[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
        for (int k = 0; k < 4; k++)
        {
          const float x = __fmul_rn(v[k], inv);
          v[k] = (isnan(x) || isinf(x)) ? 0.0f : x; // Fusion.h:79-95
        }
        put_row<VW>(dst, v, 4 * j, C);
      }
    }
    fence_proxy_async_smem();
    __syncwarp();
    float* gdst = out + (size_t) blk * 32 * C;
    const uint32_t obytes = (uint32_t) n * (uint32_t) C * 4u;
    if ((obytes & 15u) == 0u) // always for a full block (128 C bytes); the last block may not be
    {
      if (lane == 0)
      {
        bulk_s2g(gdst, out_buf[s], obytes);
        bulk_commit_group();
      }
    }
    else
    {
      for (int e = lane; e < n * C; e += 32)
      {
        gdst[e] = out_buf[s][e];
      }
      __syncwarp();
    }
    if (lane == 0 && blk + 2 * stride < nblocks)
    {
      fetch(blk + 2 * stride, s); // every lane has finished with in_buf[s] (the __syncwarp above)
    }
    if (s == 1) parity ^= 1u;
    s ^= 1;
  }
  if (lane == 0)
  {
    bulk_wait_group<0>(); // shared memory must outlive the copies that read it
  }
}

// class vectors too wide for the staged kernel: one thread per row straight from global memory
template <int KIND>
__global__ void __launch_bounds__(256) get_direct_kernel(const float* __restrict__ acc, int64_t P, int C, int Cpad,
                                                         float* __restrict__ out)
{
  const int64_t r = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= P)
  {
    return;
  }
  const float* row = acc + (size_t) r * Cpad;
  float* o = out + (size_t) r * C;
  for (int c = 0; c < C; c++)
  {
    o[c] = row[c];
  }
  get_row<KIND>(o, C);
}

template <int KIND>
static int launch_get(const float* acc, int64_t P, int C, float* out, cudaStream_t stream)
{
  const int Cpad = smesh_fuse_padded_classes(C);
  static const bool no_stream = getenv("SMESH_GET_STAGED") != nullptr; // profiling only: the staged kernel below
  // (mul: the exp() per element makes a lane's row walk the bottleneck once wide rows leave only a few warps per SM;
  // measured at C = 40 / 150: 60 / 1264 us against 46 / 1080 us of the staged kernel, at C = 19 79 against 88)
  const int stream_max_c = KIND == SMESH_KIND_MUL ? 24 : 200;
  if (C <= stream_max_c && !no_stream && (reinterpret_cast<uintptr_t>(out) & 15) == 0 && (reinterpret_cast<uintptr_t>(acc) & 15) == 0)
  {
    // warps per CTA so that a CTA stays below ~100 KB (two or more CTAs per SM): 4 up to C = 48, 2 up to 100, else 1
    const size_t warp_bytes = (size_t) 2 * 32 * (Cpad + C) * 4;
    const int warps = warp_bytes * 4 <= 100 * 1024 ? 4 : (warp_bytes * 2 <= 100 * 1024 ? 2 : 1);
    const size_t smem = warps * warp_bytes + (size_t) warps * 16;
    void (*kernel)(const float*, int64_t, int, int, float*) =
      C % 4 == 0 ? get_stream_kernel<KIND, 4> : (C % 2 == 0 ? get_stream_kernel<KIND, 2> : get_stream_kernel<KIND, 1>);
    int blocks_per_sm = 0;
    const int rc = kernel_blocks_per_sm(reinterpret_cast<const void*>(kernel), warps * 32, smem, &blocks_per_sm);
    if (rc != SMESH_OK)
    {
      return rc;
    }
    if (blocks_per_sm >= 1)
    {
      int64_t blocks = ((P + 31) / 32 + warps - 1) / warps;
      const int64_t cap = (int64_t) num_sms() * blocks_per_sm;
      if (blocks > cap) blocks = cap;
      if (blocks < 1) return SMESH_OK;
      kernel<<<(unsigned) blocks, warps * 32, smem, stream>>>(acc, P, C, Cpad, out);
      SMESH_LAUNCH_CHECK("get_stream_kernel");
      return SMESH_OK;
    }
  }
  // rows per CTA: 256 if they fit in ~96 KB (two CTAs per SM), else whatever fits in 200 KB, in whole warps
  const size_t row_bytes = (size_t) (Cpad + 1) * 4;
  static const int env_rows = getenv("SMESH_GET_ROWS") ? atoi(getenv("SMESH_GET_ROWS")) : 0; // profiling only
  int rows = (env_rows >= 32 && env_rows <= 1024) ? env_rows / 32 * 32 : (Cpad >= 32 ? 128 : 256); // (C = 40: 35 us against 39)
  if (rows * row_bytes > 96 * 1024)
  {
    rows = (int) ((200 * 1024) / row_bytes) / 32 * 32;
  }
  if (rows < 32)
  {
    get_direct_kernel<KIND><<<(unsigned) ((P + 255) / 256), 256, 0, stream>>>(acc, P, C, Cpad, out);
    SMESH_LAUNCH_CHECK("get_direct_kernel");
    return SMESH_OK;
  }
  const size_t smem = (size_t) rows * row_bytes;
  auto kernel = get_kernel<KIND>;
  int blocks_per_sm = 0;
  const int rc = kernel_blocks_per_sm(reinterpret_cast<const void*>(kernel), 256, smem, &blocks_per_sm);
  if (rc != SMESH_OK)
  {
    return rc;
  }
  int64_t blocks = (P + rows - 1) / rows;
  const int64_t cap = (int64_t) num_sms() * (blocks_per_sm > 0 ? blocks_per_sm : 1);
  if (blocks > cap) blocks = cap;
  kernel<<<(unsigned) blocks, 256, smem, stream>>>(acc, P, C, Cpad, rows, out);
  SMESH_LAUNCH_CHECK("get_kernel");
  return SMESH_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------------

struct RingConfig
{
  int consumer_warps;
  int stages;
};

static bool ring_config(int C, RingConfig& cfg)
{
  // consumer warps x stages; one stage = consumer_warps * 32 pixels * C floats (19.5 KB at C = 19)
  if (C <= 24) { cfg = {8, 3}; }
  else if (C <= 48) { cfg = {4, 2}; } // cfg2 (C = 40): 2 stages 18.2 us, 3 stages 18.8, 4 stages 20.1 (more CTAs per SM win)
  else if (C <= 96) { cfg = {4, 3}; }
  else if (C <= 192) { cfg = {4, 2}; }
  else if (C <= 400) { cfg = {2, 2}; }
  else if (C <= 800) { cfg = {1, 2}; }
  else { return false; }
  // tuning overrides (profiling only)
  static const int env_nw = getenv("SMESH_SCATTER_NW") ? atoi(getenv("SMESH_SCATTER_NW")) : 0;
  static const int env_stages = getenv("SMESH_SCATTER_STAGES") ? atoi(getenv("SMESH_SCATTER_STAGES")) : 0;
  if (env_nw >= 1 && env_nw <= 8) cfg.consumer_warps = env_nw;
  if (env_stages >= 2 && env_stages <= 8) cfg.stages = env_stages;
  return true;
}

static uint32_t scatter_run_cap()
{
  // longest run of equal face ids folded into one reduction: fewer shuffle rounds (log2) vs more reductions
  static const int env = getenv("SMESH_SCATTER_RUNCAP") ? atoi(getenv("SMESH_SCATTER_RUNCAP")) : 0;
  return (env == 1 || env == 2 || env == 4 || env == 8 || env == 16 || env == 32) ? (uint32_t) env : 32u;
}

static int ring_pad_group_enabled(int C)
{
  static const bool off = getenv("SMESH_NO_RING_PAD") != nullptr; // profiling only
  return off ? 0 : ring_pad_group(C);
}

static size_t ring_smem_bytes(int C, const RingConfig& cfg)
{
  const int padg = ring_pad_group_enabled(C);
  return (size_t) cfg.stages * (cfg.consumer_warps * 32 * C * 4 + (padg > 0 ? (cfg.consumer_warps * 32 / padg) * 16 : 0)) +
         (size_t) cfg.stages * 16;
}

template <int KIND, int CT>
static int launch_scatter_ring(const ScatterArgs& args_in, const RingConfig& cfg, cudaStream_t stream)
{
  ScatterArgs args = args_in;
  const size_t smem = ring_smem_bytes(args.C, cfg);
  const bool rider = args.next_ids != nullptr;
  // (a compile-time C knows its padding; the run-time instance carries the builds for 0 / 2 / 4)
  const int padg = ring_pad_group_enabled(args.C);
  void (*kernel)(ScatterArgs) = nullptr;
  if (CT > 0)
  {
    constexpr int PG = CT > 0 ? ring_pad_group(CT) : 0;
    if (padg == PG) kernel = rider ? scatter_kernel<KIND, CT, true, PG> : scatter_kernel<KIND, CT, false, PG>;
    else kernel = rider ? scatter_kernel<KIND, CT, true, 0> : scatter_kernel<KIND, CT, false, 0>;
  }
  else if (padg == 4) kernel = rider ? scatter_kernel<KIND, CT, true, 4> : scatter_kernel<KIND, CT, false, 4>;
  else if (padg == 2) kernel = rider ? scatter_kernel<KIND, CT, true, 2> : scatter_kernel<KIND, CT, false, 2>;
  else kernel = rider ? scatter_kernel<KIND, CT, true, 0> : scatter_kernel<KIND, CT, false, 0>;
  const int threads = (cfg.consumer_warps + 1 + (rider ? 1 : 0)) * 32;
  int blocks_per_sm = 0;
  const int cfg_rc = kernel_blocks_per_sm(reinterpret_cast<const void*>(kernel), threads, smem, &blocks_per_sm);
  if (cfg_rc != SMESH_OK)
  {
    return cfg_rc;
  }
  if (blocks_per_sm < 1)
  {
    set_error("scatter_kernel does not fit on an SM (C=%d, %zu bytes of shared memory)", args.C, smem);
    return SMESH_ERR_UNSUPPORTED;
  }
  const int tile_px = cfg.consumer_warps * 32;
  args.ntiles = (args.npix + tile_px - 1) / tile_px;
  args.stages = cfg.stages;
  int64_t blocks = (int64_t) num_sms() * blocks_per_sm;
  if (blocks > args.ntiles) blocks = args.ntiles;
  if (blocks < 1) return SMESH_OK;
  kernel<<<(unsigned) blocks, threads, smem, stream>>>(args);
  SMESH_LAUNCH_CHECK("scatter_kernel");
  return SMESH_OK;
}

struct PairConfig
{
  int consumer_warps;
  int stages;
};

static PairConfig pair_config(int C)
{
  // 3 consumer warps x 2 stages = 45.5 KB per CTA, four CTAs per SM: the same speed alone as 4 warps x 3 CTAs, and 7 % more
  // views/s when the rasterizer shares the SMs (finer-grained CTAs interleave better), measured on cfg3 (C = 19).
  // Narrower class vectors get more stages so that a CTA keeps about the same 28 KB of bulk copies in flight.
  PairConfig cfg = {3, 2};
  const int stage_bytes = cfg.consumer_warps * 64 * C * 4;
  cfg.stages = std::min(8, std::max(2, (28 * 1024 + stage_bytes - 1) / stage_bytes));
  static const int env_nw = getenv("SMESH_PAIR_NW") ? atoi(getenv("SMESH_PAIR_NW")) : 0;
  static const int env_stages = getenv("SMESH_PAIR_STAGES") ? atoi(getenv("SMESH_PAIR_STAGES")) : 0;
  if (env_nw >= 1 && env_nw <= 8) cfg.consumer_warps = env_nw;
  if (env_stages >= 2 && env_stages <= 8) cfg.stages = env_stages;
  return cfg;
}

// Which class counts take the padded layout: measured (profiles/r02s_pair_even_class_counts.txt, cfg3's index images,
// launches back to back): C = 12: 30.1 -> 26.8 us, C = 20: 40.7 -> 39.0 us. Not taken: C = 4, 8, 16 - their groups are bulk
// copies of 128 bytes, and that many small copies cost more than the conflicts (17.5 -> 22.7, 26.4 -> 38.3, 58.2 -> 66.8 us);
// C = 2 (mod 4): 128-bit loads alone change nothing (the kernel is not bound by the conflicts there).
constexpr bool pair_wide_default(int C)
{
  return C == 20 || C == 12;
}

template <int KIND, int CT>
static int launch_scatter_pair(const ScatterArgs& args_in, cudaStream_t stream)
{
  ScatterArgs args = args_in;
  const PairConfig cfg = pair_config(CT);
  // C = 12, 20: the padded layout (see scatter_pair_kernel); SMESH_PAIR_WIDE=0 switches it off (tuning)
  static const int env_wide = getenv("SMESH_PAIR_WIDE") ? atoi(getenv("SMESH_PAIR_WIDE")) : -1;
  const bool wide = pair_wide_default(CT) && env_wide != 0;
  const int padl = wide ? pair_pad_lanes(CT) : 0;
  const size_t smem = (size_t) cfg.stages * (cfg.consumer_warps * 64 * CT * 4 + (padl > 0 ? (cfg.consumer_warps * 32 / padl) * 16 : 0)) +
                      (size_t) cfg.stages * 16 +
                      (size_t) cfg.consumer_warps * (64 * ((CT + 3) & ~3) + 64) * 4; // flush rows + their face ids
  void (*kernel)(ScatterArgs) = scatter_pair_kernel<KIND, CT, false>;
  if constexpr (pair_wide_default(CT))
  {
    if (wide) kernel = scatter_pair_kernel<KIND, CT, true>;
  }
  const int threads = (cfg.consumer_warps + 1 + (args.next_ids != nullptr ? 1 : 0)) * 32;
  int blocks_per_sm = 0;
  const int cfg_rc = kernel_blocks_per_sm(reinterpret_cast<const void*>(kernel), threads, smem, &blocks_per_sm);
  if (cfg_rc != SMESH_OK)
  {
    return cfg_rc;
  }
  if (blocks_per_sm < 1)
  {
    set_error("scatter_pair_kernel does not fit on an SM (%zu bytes of shared memory)", smem);
    return SMESH_ERR_UNSUPPORTED;
  }
  const int tile_px = cfg.consumer_warps * 64;
  args.ntiles = (args.npix + tile_px - 1) / tile_px;
  args.stages = cfg.stages;
  static const int env_ctas = getenv("SMESH_PAIR_CTAS") ? atoi(getenv("SMESH_PAIR_CTAS")) : 4; // see pair_config()
  int64_t blocks = (int64_t) num_sms() * (env_ctas >= 1 && env_ctas < blocks_per_sm ? env_ctas : blocks_per_sm);
  if (blocks > args.ntiles) blocks = args.ntiles;
  if (blocks < 1) return SMESH_OK;
  kernel<<<(unsigned) blocks, threads, smem, stream>>>(args);
  SMESH_LAUNCH_CHECK("scatter_pair_kernel");
  return SMESH_OK;
}

// Can the scatter launch of these arguments carry the next view's count stage? (the two ring kernels can)
static bool scatter_takes_count_job(int kind, const ScatterArgs& args);

template <int KIND>
static int launch_scatter(const ScatterArgs& args, cudaStream_t stream)
{
  RingConfig cfg;
  const bool aligned = (reinterpret_cast<uintptr_t>(args.probs) & 15) == 0;
  static const bool no_pair = getenv("SMESH_NO_PAIR") != nullptr;
  if (KIND != SMESH_KIND_SUMMAX)
  {
    int vw = 1;
    if (rows_kernel_takes(args, vw))
    {
      constexpr int K = KIND == SMESH_KIND_SUMMAX ? SMESH_KIND_SUM : KIND;
      return vw == 4 ? launch_scatter_rows<K, 4>(args, stream)
                     : (vw == 2 ? launch_scatter_rows<K, 2>(args, stream) : launch_scatter_rows<K, 1>(args, stream));
    }
  }
  if (aligned && ring_config(args.C, cfg))
  {
    if (KIND != SMESH_KIND_SUMMAX && !no_pair && (reinterpret_cast<uintptr_t>(args.ids) & 7) == 0)
    {
      // narrow class vectors (2 ... CH classes): two pixels per lane
      constexpr int K = KIND == SMESH_KIND_SUMMAX ? SMESH_KIND_SUM : KIND;
      switch (args.C)
      {
#define SMESH_PAIR_CASE(c) case c: return launch_scatter_pair<K, c>(args, stream);
        SMESH_PAIR_CASE(2) SMESH_PAIR_CASE(3) SMESH_PAIR_CASE(4) SMESH_PAIR_CASE(5) SMESH_PAIR_CASE(6)
        SMESH_PAIR_CASE(7) SMESH_PAIR_CASE(8) SMESH_PAIR_CASE(9) SMESH_PAIR_CASE(10) SMESH_PAIR_CASE(11)
        SMESH_PAIR_CASE(12) SMESH_PAIR_CASE(13) SMESH_PAIR_CASE(14) SMESH_PAIR_CASE(15) SMESH_PAIR_CASE(16)
        SMESH_PAIR_CASE(17) SMESH_PAIR_CASE(18) SMESH_PAIR_CASE(19) SMESH_PAIR_CASE(20)
#undef SMESH_PAIR_CASE
        default: break;
      }
    }
    switch (args.C)
    {
      case 19: return launch_scatter_ring<KIND, 19>(args, cfg, stream);
      case 40: return launch_scatter_ring<KIND, 40>(args, cfg, stream);
      default: return launch_scatter_ring<KIND, 0>(args, cfg, stream);
    }
  }
  const int64_t blocks = (args.npix + 255) / 256;
  if (blocks > 0)
  {
    scatter_direct_kernel<KIND><<<(unsigned) blocks, 256, 0, stream>>>(args);
    SMESH_LAUNCH_CHECK("scatter_direct_kernel");
  }
  return SMESH_OK;
}

template <typename IdT>
static int launch_count(const void* ids, int64_t so, int64_t si, int64_t n_outer, int64_t n_inner, int64_t P, uint32_t* counts,
                        uint32_t* ids32, uint32_t epoch, cudaStream_t stream)
{
  const int64_t npix = n_outer * n_inner;
  const bool flat = (si == 1 || n_inner == 1) && (so == n_inner || n_outer == 1);
  const int64_t per_block = (256 / 32) * 32 * COUNT_UNROLL;
  const int64_t blocks = std::min((npix + per_block - 1) / per_block, (int64_t) num_sms() * 8);
  count_runs_kernel<IdT, COUNT_UNROLL><<<(unsigned) blocks, 256, 0, stream>>>(static_cast<const IdT*>(ids), so, si, n_inner, npix, P,
                                                                             counts, ids32, flat ? 1 : 0, epoch);
  SMESH_LAUNCH_CHECK("count_runs_kernel");
  return SMESH_OK;
}

static int launch_count_any(const char* fn, int id_dtype, const void* ids, int64_t so, int64_t si, int64_t n_outer,
                            int64_t n_inner, int64_t P, uint32_t* counts, uint32_t* ids32, uint32_t epoch, cudaStream_t stream)
{
  switch (id_dtype)
  {
    case SMESH_ID_U32: return launch_count<uint32_t>(ids, so, si, n_outer, n_inner, P, counts, ids32, epoch, stream);
    case SMESH_ID_I32: return launch_count<int32_t>(ids, so, si, n_outer, n_inner, P, counts, ids32, epoch, stream);
    case SMESH_ID_U64: return launch_count<uint64_t>(ids, so, si, n_outer, n_inner, P, counts, ids32, epoch, stream);
    case SMESH_ID_I64: return launch_count<int64_t>(ids, so, si, n_outer, n_inner, P, counts, ids32, epoch, stream);
    default: set_error("%s: unknown id dtype %d", fn, id_dtype); return SMESH_ERR_INVALID_ARGUMENT;
  }
}

static int check_epoch(const char* fn, uint32_t epoch, int64_t npix)
{
  if (epoch > 255u)
  {
    set_error("%s: count_epoch %u out of range (0..255)", fn, epoch);
    return SMESH_ERR_INVALID_ARGUMENT;
  }
  if (epoch != 0 && npix > (int64_t) COUNT_MASK)
  {
    set_error("%s: images of %lld pixels need count_epoch 0 (tagged counters hold 24 bits)", fn, (long long) npix);
    return SMESH_ERR_UNSUPPORTED;
  }
  return SMESH_OK;
}

static ScatterArgs make_scatter_args(const uint32_t* ids32, const float* probs, const float* weights, const uint32_t* counts,
                                     float* acc, int64_t npix, int C, int64_t P, float iew, uint32_t epoch)
{
  ScatterArgs args;
  args.probs = probs;
  args.ids = ids32;
  args.weights = weights;
  args.counts = counts;
  args.acc = acc;
  args.npix = npix;
  args.P = P;
  args.ntiles = 0;
  args.C = C;
  args.Cpad = smesh_fuse_padded_classes(C);
  args.stages = 0;
  args.count_mask = epoch != 0 ? COUNT_MASK : 0xFFFFFFFFu;
  args.run_cap = scatter_run_cap();
  const char* mul_exact_env = getenv("SMESH_MUL_EXACT"); // read per call: tests switch it
  args.mul_exact = (mul_exact_env != nullptr && atoi(mul_exact_env) != 0) ? 1 : 0;
  args.iew = iew;
  static const uint32_t wait_hint = getenv("SMESH_WAIT_HINT") ? (uint32_t) strtoul(getenv("SMESH_WAIT_HINT"), nullptr, 0) : 0x989680u;
  args.wait_hint = wait_hint;
  args.next_ids = nullptr;
  args.next_counts = nullptr;
  args.next_npix = 0;
  args.next_tag = 0;
  return args;
}

static bool scatter_takes_count_job(int kind, const ScatterArgs& args)
{
  // mirrors launch_scatter(): class-parallel kernel for wide C (no spare warp), else a ring kernel if the image is aligned.
  // summax: its scatter is one scalar reduction per run - with the count warp beside it the launch was measured slower
  // than the two stages one after the other (cfg3: 46.8 vs 38.9 us per view), so it takes none.
  int vw = 1;
  RingConfig cfg;
  if (kind == SMESH_KIND_SUMMAX || rows_kernel_takes(args, vw))
  {
    return false;
  }
  return (reinterpret_cast<uintptr_t>(args.probs) & 15) == 0 && ring_config(args.C, cfg);
}

static int launch_scatter_kind(int kind, const ScatterArgs& args, cudaStream_t stream)
{
  switch (kind)
  {
    case SMESH_KIND_SUM: return launch_scatter<SMESH_KIND_SUM>(args, stream);
    case SMESH_KIND_SUMMAX: return launch_scatter<SMESH_KIND_SUMMAX>(args, stream);
    case SMESH_KIND_MUL: return launch_scatter<SMESH_KIND_MUL>(args, stream);
    default: set_error("unknown aggregator kind %d", kind); return SMESH_ERR_INVALID_ARGUMENT;
  }
}

// One view = two stages on possibly different streams: the per-face pixel count (Mesh.h:90-93) and the gated, weighted
// scatter (Mesh.h:94-106).
struct ViewStages
{
  int kind, id_dtype, C;
  const void* ids;
  int64_t ids_so, ids_si, n_outer, n_inner, P;
  const float* probs;
  const float* weights;
  float iew;
  uint32_t* ids32;
  float* acc;
  bool zero_copy; // 32-bit ids already in flat order are consumed in place

  int64_t npix() const { return n_outer * n_inner; }
  const uint32_t* flat_ids() const { return zero_copy ? static_cast<const uint32_t*>(ids) : ids32; }

  int count(uint32_t* counts, uint32_t epoch, cudaStream_t stream) const
  {
    return launch_count_any("smesh_fuse_add", id_dtype, ids, ids_so, ids_si, n_outer, n_inner, P, counts,
                            zero_copy ? nullptr : ids32, epoch, stream);
  }

  // next / next_counts / next_epoch: the view whose count stage rides along (NULL = none); see ScatterArgs
  int scatter(uint32_t* counts, uint32_t epoch, cudaStream_t stream, const ViewStages* next = nullptr,
              uint32_t* next_counts = nullptr, uint32_t next_epoch = 0) const
  {
    ScatterArgs args = make_scatter_args(flat_ids(), probs, weights, counts, acc, npix(), C, P, iew, epoch);
    if (next != nullptr)
    {
      args.next_ids = next->flat_ids();
      args.next_counts = next_counts;
      args.next_npix = next->npix();
      args.next_tag = next_epoch << COUNT_BITS;
    }
    const int rc = launch_scatter_kind(kind, args, stream);
    if (rc != SMESH_OK)
    {
      return rc;
    }
    if (epoch == 0)
    {
      clear_kernel<<<(unsigned) ((npix() + 255) / 256), 256, 0, stream>>>(flat_ids(), npix(), P, counts);
      SMESH_LAUNCH_CHECK("clear_kernel");
    }
    return SMESH_OK;
  }
};

// The side stream of smesh_fuse_add_batch's second lane and its fork / join events: one set per host thread and device
// (two host threads batching on one device must not share events). NULL while the caller's stream is being captured and
// the set does not exist yet (nothing is created during a capture), or if creation fails: the caller then takes the
// single-stream path.
struct BatchLane
{
  cudaStream_t side = nullptr;
  cudaEvent_t fork = nullptr, join = nullptr;
};

static BatchLane* batch_lane(cudaStream_t stream)
{
  constexpr int MAX_DEVICES = 64;
  thread_local BatchLane lanes[MAX_DEVICES];
  int dev = -1;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= MAX_DEVICES)
  {
    cudaGetLastError();
    return nullptr;
  }
  BatchLane& l = lanes[dev];
  if (l.side == nullptr)
  {
    cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(stream, &st) != cudaSuccess || st != cudaStreamCaptureStatusNone)
    {
      cudaGetLastError();
      return nullptr;
    }
    BatchLane fresh;
    if (cudaStreamCreateWithFlags(&fresh.side, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&fresh.fork, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&fresh.join, cudaEventDisableTiming) != cudaSuccess)
    {
      cudaGetLastError();
      if (fresh.join) cudaEventDestroy(fresh.join);
      if (fresh.fork) cudaEventDestroy(fresh.fork);
      if (fresh.side) cudaStreamDestroy(fresh.side);
      return nullptr;
    }
    l = fresh;
  }
  return &l;
}

static int make_view(const char* fn, ViewStages& v, int kind, const void* ids, int id_dtype, int64_t ids_so, int64_t ids_si,
                     const float* probs, const float* weights, int64_t w_so, int64_t w_si, int64_t n_outer, int64_t n_inner,
                     int C, int64_t P, float iew, uint32_t* ids32, float* acc, uint32_t epoch)
{
  const int rc = check_epoch(fn, epoch, n_outer * n_inner);
  if (rc != SMESH_OK)
  {
    return rc;
  }
  if (weights != nullptr)
  {
    const bool w_flat = (w_si == 1 || n_inner == 1) && (w_so == n_inner || n_outer == 1);
    if (!w_flat)
    {
      set_error("%s: the weights image must be contiguous in the same pixel order as the probability image", fn);
      return SMESH_ERR_UNSUPPORTED;
    }
  }
  const bool ids_flat = (ids_si == 1 || n_inner == 1) && (ids_so == n_inner || n_outer == 1);
  v.kind = kind; v.id_dtype = id_dtype; v.C = C;
  v.ids = ids; v.ids_so = ids_so; v.ids_si = ids_si; v.n_outer = n_outer; v.n_inner = n_inner; v.P = P;
  v.probs = probs; v.weights = weights; v.iew = iew; v.ids32 = ids32; v.acc = acc;
  // (int32: negative values read as >= 2^31 and fail `id < P`)
  v.zero_copy = ids_flat && (id_dtype == SMESH_ID_U32 || (id_dtype == SMESH_ID_I32 && P <= 0x7FFFFFFFll));
  return SMESH_OK;
}

static int add_view(int kind, const void* ids, int id_dtype, int64_t ids_so, int64_t ids_si, const float* probs,
                    const float* weights, int64_t w_so, int64_t w_si, int64_t n_outer, int64_t n_inner, int C, int64_t P,
                    float iew, uint32_t* counts, uint32_t* ids32, float* acc, uint32_t epoch, cudaStream_t stream)
{
  if (n_outer * n_inner == 0 || P == 0)
  {
    return SMESH_OK;
  }
  ViewStages v;
  int rc = make_view("smesh_fuse_add", v, kind, ids, id_dtype, ids_so, ids_si, probs, weights, w_so, w_si, n_outer, n_inner, C, P,
                     iew, ids32, acc, epoch);
  if (rc != SMESH_OK)
  {
    return rc;
  }
  rc = v.count(counts, epoch, stream);
  return rc != SMESH_OK ? rc : v.scatter(counts, epoch, stream);
}

static int check_add_args(const char* fn, int kind, const void* ids, const float* probs, int64_t n_outer, int64_t n_inner,
                          int C, int64_t P, const uint32_t* counts, const uint32_t* ids32, const float* acc)
{
  if (n_outer < 0 || n_inner < 0 || C < 1 || P < 0 || kind < 0 || kind > 2)
  {
    set_error("%s: invalid argument (n_outer=%lld n_inner=%lld C=%d P=%lld kind=%d)", fn, (long long) n_outer,
              (long long) n_inner, C, (long long) P, kind);
    return SMESH_ERR_INVALID_ARGUMENT;
  }
  if (n_outer * n_inner > 0 && P > 0 && (!ids || !probs || !counts || !ids32 || !acc))
  {
    set_error("%s: null buffer", fn);
    return SMESH_ERR_INVALID_ARGUMENT;
  }
  if (C > 4096 || P >= 0xFFFFFFFFll)
  {
    set_error("%s: unsupported size (C=%d must be <= 4096, P=%lld < 2^32-1)", fn, C, (long long) P);
    return SMESH_ERR_UNSUPPORTED;
  }
  return SMESH_OK;
}

} // namespace fuse
} // namespace smesh

using namespace smesh;
using namespace smesh::fuse;

extern "C" int smesh_fuse_padded_classes(int C)
{
  return (C + 3) & ~3;
}

extern "C" int smesh_fuse_add(int kind, const void* ids, int id_dtype, int64_t ids_stride_outer, int64_t ids_stride_inner,
                              const float* probs, const float* weights, int64_t w_stride_outer, int64_t w_stride_inner,
                              int64_t n_outer, int64_t n_inner, int C, int64_t P, float iew, uint32_t* counts,
                              uint32_t count_epoch, uint32_t* ids32, float* acc, void* stream)
{
  const int rc = check_add_args("smesh_fuse_add", kind, ids, probs, n_outer, n_inner, C, P, counts, ids32, acc);
  if (rc != SMESH_OK)
  {
    return rc;
  }
  return add_view(kind, ids, id_dtype, ids_stride_outer, ids_stride_inner, probs, weights, w_stride_outer, w_stride_inner,
                  n_outer, n_inner, C, P, iew, counts, ids32, acc, count_epoch, static_cast<cudaStream_t>(stream));
}

extern "C" int smesh_fuse_count(const void* ids, int id_dtype, int64_t ids_stride_outer, int64_t ids_stride_inner,
                                int64_t n_outer, int64_t n_inner, int64_t P, uint32_t* counts, uint32_t count_epoch,
                                uint32_t* ids32_out, void* stream_v)
{
  if (n_outer < 0 || n_inner < 0 || P < 0 || P >= 0xFFFFFFFFll || (n_outer * n_inner > 0 && P > 0 && (!ids || !counts)))
  {
    set_error("smesh_fuse_count: invalid argument");
    return SMESH_ERR_INVALID_ARGUMENT;
  }
  const int64_t npix = n_outer * n_inner;
  if (npix == 0 || P == 0)
  {
    return SMESH_OK;
  }
  const int rc = check_epoch("smesh_fuse_count", count_epoch, npix);
  if (rc != SMESH_OK)
  {
    return rc;
  }
  return launch_count_any("smesh_fuse_count", id_dtype, ids, ids_stride_outer, ids_stride_inner, n_outer, n_inner, P, counts,
                          ids32_out, count_epoch, static_cast<cudaStream_t>(stream_v));
}

extern "C" int smesh_fuse_scatter(int kind, const uint32_t* ids32, const float* probs, const float* weights, int64_t n_pix,
                                  int C, int64_t P, float iew, const uint32_t* counts, uint32_t count_epoch, float* acc,
                                  void* stream_v)
{
  if (n_pix < 0 || C < 1 || P < 0 || kind < 0 || kind > 2 || (n_pix > 0 && P > 0 && (!ids32 || !probs || !counts || !acc)))
  {
    set_error("smesh_fuse_scatter: invalid argument");
    return SMESH_ERR_INVALID_ARGUMENT;
  }
  if (C > 4096 || P >= 0xFFFFFFFFll)
  {
    set_error("smesh_fuse_scatter: unsupported size (C=%d must be <= 4096, P=%lld < 2^32-1)", C, (long long) P);
    return SMESH_ERR_UNSUPPORTED;
  }
  if (n_pix == 0 || P == 0)
  {
    return SMESH_OK;
  }
  const int rc = check_epoch("smesh_fuse_scatter", count_epoch, n_pix);
  if (rc != SMESH_OK)
  {
    return rc;
  }
  return launch_scatter_kind(kind, make_scatter_args(ids32, probs, weights, counts, acc, n_pix, C, P, iew, count_epoch),
                             static_cast<cudaStream_t>(stream_v));
}

extern "C" int smesh_fuse_scatter_count_next(int kind, const uint32_t* ids32, const float* probs, const float* weights,
                                            int64_t n_pix, int C, int64_t P, float iew, uint32_t* counts, uint32_t count_epoch,
                                            int counted, const uint32_t* next_ids32, int64_t next_n_pix, uint32_t* next_counts,
                                            uint32_t next_epoch, float* acc, void* stream_v)
{
  if (n_pix < 0 || next_n_pix < 0 || C < 1 || P < 0 || kind < 0 || kind > 2 ||
      (n_pix > 0 && P > 0 && (!ids32 || !probs || !counts || !acc)) || (next_n_pix > 0 && P > 0 && (!next_ids32 || !next_counts)))
  {
    set_error("smesh_fuse_scatter_count_next: invalid argument");
    return SMESH_ERR_INVALID_ARGUMENT;
  }
  if (C > 4096 || P >= 0xFFFFFFFFll)
  {
    set_error("smesh_fuse_scatter_count_next: unsupported size (C=%d must be <= 4096, P=%lld < 2^32-1)", C, (long long) P);
    return SMESH_ERR_UNSUPPORTED;
  }
  if (count_epoch == 0 || next_epoch == 0 || count_epoch > 255u || next_epoch > 255u || ((count_epoch ^ next_epoch) & 1u) == 0u ||
      next_counts == counts || n_pix > (int64_t) COUNT_MASK || next_n_pix > (int64_t) COUNT_MASK)
  {
    set_error("smesh_fuse_scatter_count_next: needs tagged epochs 1..255 of different parity in two counter arrays");
    return SMESH_ERR_INVALID_ARGUMENT;
  }
  if (P == 0)
  {
    return SMESH_OK;
  }
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  int rc = SMESH_OK;
  if (!counted && n_pix > 0)
  {
    rc = launch_count_any("smesh_fuse_scatter_count_next", SMESH_ID_U32, ids32, n_pix, 1, 1, n_pix, P, counts, nullptr, count_epoch,
                          stream);
    if (rc != SMESH_OK)
    {
      return rc;
    }
  }
  bool carried = false;
  if (n_pix > 0)
  {
    ScatterArgs args = make_scatter_args(ids32, probs, weights, counts, acc, n_pix, C, P, iew, count_epoch);
    static const bool no_overlap = getenv("SMESH_NO_BATCH_OVERLAP") != nullptr; // profiling only
    if (next_n_pix > 0 && !no_overlap && scatter_takes_count_job(kind, args))
    {
      args.next_ids = next_ids32;
      args.next_counts = next_counts;
      args.next_npix = next_n_pix;
      args.next_tag = next_epoch << COUNT_BITS;
      carried = true;
    }
    rc = launch_scatter_kind(kind, args, stream);
    if (rc != SMESH_OK)
    {
      return rc;
    }
  }
  if (!carried && next_n_pix > 0)
  {
    // this view's scatter kernel has no spare warp: the next view's count stage is a launch of its own
    rc = launch_count_any("smesh_fuse_scatter_count_next", SMESH_ID_U32, next_ids32, next_n_pix, 1, 1, next_n_pix, P, next_counts,
                          nullptr, next_epoch, stream);
  }
  return rc;
}

extern "C" int smesh_fuse_clear(const uint32_t* ids32, int64_t n_pix, int64_t P, uint32_t* counts, void* stream_v)
{
  if (n_pix < 0 || P < 0 || (n_pix > 0 && P > 0 && (!ids32 || !counts)))
  {
    set_error("smesh_fuse_clear: invalid argument");
    return SMESH_ERR_INVALID_ARGUMENT;
  }
  if (n_pix == 0 || P == 0)
  {
    return SMESH_OK;
  }
  clear_kernel<<<(unsigned) ((n_pix + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream_v)>>>(ids32, n_pix, P, counts);
  SMESH_LAUNCH_CHECK("clear_kernel");
  return SMESH_OK;
}

extern "C" int smesh_fuse_add_batch(int kind, int64_t B, const void* ids, int id_dtype, int64_t ids_stride_view,
                                    int64_t ids_stride_outer, int64_t ids_stride_inner, const float* probs,
                                    int64_t probs_stride_view, const float* weights, int64_t w_stride_view,
                                    int64_t w_stride_outer, int64_t w_stride_inner, int64_t n_outer, int64_t n_inner, int C,
                                    int64_t P, float iew, uint32_t* counts2, uint32_t count_epoch0, uint32_t* ids32, float* acc,
                                    void* stream_v)
{
  int rc = check_add_args("smesh_fuse_add_batch", kind, ids, probs, n_outer, n_inner, C, P, counts2, ids32, acc);
  if (rc != SMESH_OK)
  {
    return rc;
  }
  if (B < 0 || (count_epoch0 != 0 && (int64_t) count_epoch0 + B - 1 > 255))
  {
    set_error("smesh_fuse_add_batch: invalid batch (B=%lld, count_epoch0=%u: epochs must stay <= 255)", (long long) B,
              count_epoch0);
    return SMESH_ERR_INVALID_ARGUMENT;
  }
  if (B == 0 || n_outer * n_inner == 0 || P == 0)
  {
    return SMESH_OK;
  }
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  const size_t id_size = (id_dtype == SMESH_ID_U64 || id_dtype == SMESH_ID_I64) ? 8 : 4;
  auto view = [&](int64_t b, ViewStages& v) -> int {
    const void* ids_b = static_cast<const char*>(ids) + (size_t) b * ids_stride_view * id_size;
    const float* probs_b = probs + (size_t) b * probs_stride_view;
    const float* weights_b = weights ? weights + (size_t) b * w_stride_view : nullptr;
    return make_view("smesh_fuse_add_batch", v, kind, ids_b, id_dtype, ids_stride_outer, ids_stride_inner, probs_b, weights_b,
                     w_stride_outer, w_stride_inner, n_outer, n_inner, C, P, iew, ids32, acc, count_epoch0);
  };
  // tagged mode: view b counts into array (epoch & 1); untagged mode: array 0 only (it is clean again after every view)
  auto epoch_of = [&](int64_t b) -> uint32_t { return count_epoch0 != 0 ? count_epoch0 + (uint32_t) b : 0u; };
  auto counts_of = [&](int64_t b) -> uint32_t* { return counts2 + (size_t) (epoch_of(b) & 1u) * (size_t) P; };

  // Two lanes: with tagged counters and ids consumed in place, the views alternate between the caller's stream and a side
  // stream of this library (forked from / joined to the caller's stream by events, so the call stays stream-ordered and
  // capturable). Each lane runs count(b) -> scatter(b) of its views on its own counter array (array = epoch parity = lane);
  // the accumulator is shared (atomic reductions commute). One lane's count stage, launch gap, ramp-up and tail are
  // covered by the other lane's scatter kernel: measured 36.3 us per view at cfg3 against 39.6 us with the riding count
  // below and 43.5 us for one add() per view (cfg5: 18.2 / 21.4 / 23.1), tools/time_two_streams.py.
  const bool no_lanes = getenv("SMESH_NO_BATCH_LANES") != nullptr; // profiling / tests: the single-stream path below
  if (B >= 2 && count_epoch0 != 0 && !no_lanes)
  {
    ViewStages probe;
    rc = view(0, probe);
    if (rc != SMESH_OK)
    {
      return rc;
    }
    BatchLane* lane = probe.zero_copy ? batch_lane(stream) : nullptr;
    if (lane != nullptr)
    {
      SMESH_CUDA_CHECK(cudaEventRecord(lane->fork, stream));
      SMESH_CUDA_CHECK(cudaStreamWaitEvent(lane->side, lane->fork, 0));
      for (int64_t b = 0; b < B; b++)
      {
        ViewStages v;
        rc = view(b, v);
        cudaStream_t st = (b & 1) ? lane->side : stream;
        if (rc == SMESH_OK) rc = v.count(counts_of(b), epoch_of(b), st);
        if (rc == SMESH_OK) rc = v.scatter(counts_of(b), epoch_of(b), st);
        if (rc != SMESH_OK)
        {
          break;
        }
      }
      // (join even after an error: the caller's stream must not run ahead of work already queued on the side stream)
      SMESH_CUDA_CHECK(cudaEventRecord(lane->join, lane->side));
      SMESH_CUDA_CHECK(cudaStreamWaitEvent(stream, lane->join, 0));
      return rc;
    }
  }

  static const bool no_overlap = getenv("SMESH_NO_BATCH_OVERLAP") != nullptr; // profiling only
  // The count stage of view b+1 rides in the scatter launch of view b (one extra warp per CTA of the ring kernels, see
  // ScatterArgs): it needs tagged counters (two arrays in flight), ids that are consumed in place (no ids32 scratch) and a
  // scatter kernel with a spare warp. Otherwise the stages simply alternate.
  bool counted = false; // view b's counts are already in flight
  for (int64_t b = 0; b < B; b++)
  {
    ViewStages v, vn;
    rc = view(b, v);
    if (rc != SMESH_OK)
    {
      return rc;
    }
    if (!counted)
    {
      rc = v.count(counts_of(b), epoch_of(b), stream);
      if (rc != SMESH_OK)
      {
        return rc;
      }
    }
    bool carry = false;
    if (b + 1 < B && count_epoch0 != 0 && !no_overlap)
    {
      rc = view(b + 1, vn);
      if (rc != SMESH_OK)
      {
        return rc;
      }
      carry = vn.zero_copy && v.zero_copy &&
              scatter_takes_count_job(kind, make_scatter_args(v.flat_ids(), v.probs, v.weights, counts_of(b), acc, v.npix(), C, P,
                                                              iew, epoch_of(b)));
    }
    rc = carry ? v.scatter(counts_of(b), epoch_of(b), stream, &vn, counts_of(b + 1), epoch_of(b + 1))
               : v.scatter(counts_of(b), epoch_of(b), stream);
    if (rc != SMESH_OK)
    {
      return rc;
    }
    counted = carry;
  }
  return SMESH_OK;
}

extern "C" int smesh_fuse_get(int kind, const float* acc, int64_t P, int C, float* out, void* stream_v)
{
  if (P < 0 || C < 1 || kind < 0 || kind > 2 || (P > 0 && (!acc || !out)))
  {
    set_error("smesh_fuse_get: invalid argument");
    return SMESH_ERR_INVALID_ARGUMENT;
  }
  if (P == 0)
  {
    return SMESH_OK;
  }
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  switch (kind)
  {
    case SMESH_KIND_SUM: return launch_get<SMESH_KIND_SUM>(acc, P, C, out, stream);
    case SMESH_KIND_SUMMAX: return launch_get<SMESH_KIND_SUMMAX>(acc, P, C, out, stream);
    default: return launch_get<SMESH_KIND_MUL>(acc, P, C, out, stream);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// What follows get() in the reference's pipeline: per-face labels and the gather-back of per-face
// annotations into an image.
// ---------------------------------------------------------------------------------------------------------------------

namespace smesh {
namespace fuse {

// python/scripts/colorize_mesh.py:82-88 on the distribution get() returns: a face whose distribution sums to less than
// the threshold received no annotation (-1); otherwise the FIRST class of maximal probability (tf.argmax).
__global__ void __launch_bounds__(256) labels_kernel(const float* __restrict__ dist, int64_t P, int C, float threshold,
                                                     int32_t* __restrict__ labels)
{
  const int64_t r = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= P)
  {
    return;
  }
  const float* row = dist + (size_t) r * C;
  float sum = 0.0f, best = row[0];
  int best_c = 0;
  for (int c = 0; c < C; c++)
  {
    const float v = row[c];
    sum = __fadd_rn(sum, v);
    if (v > best)
    {
      best = v;
      best_c = c;
    }
  }
  labels[r] = sum < threshold ? -1 : best_c;
}

// ModelRenderer::render (include/semantic_meshes/fusion/Mesh.h:24-43): dest(pixel) = annotations(primitive) if the
// primitive index is valid, else background. Elements are `words` 32-bit words (or `bytes` bytes when words == 0).
__global__ void __launch_bounds__(256) gather_kernel(const unsigned char* __restrict__ annotations, int64_t P, int words,
                                                     int bytes, const uint32_t* __restrict__ ids, int64_t npix,
                                                     const unsigned char* __restrict__ background,
                                                     unsigned char* __restrict__ out)
{
  const int64_t per = words > 0 ? words : bytes;
  const int64_t total = npix * per;
  for (int64_t t = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t) gridDim.x * blockDim.x)
  {
    const int64_t i = t / per, k = t - i * per;
    const uint32_t id = ids[i];
    const bool valid = (int64_t) id < P;
    if (words > 0)
    {
      const uint32_t* src = valid ? reinterpret_cast<const uint32_t*>(annotations) + (size_t) id * words
                                  : reinterpret_cast<const uint32_t*>(background);
      reinterpret_cast<uint32_t*>(out)[t] = src[k];
    }
    else
    {
      const unsigned char* src = valid ? annotations + (size_t) id * bytes : background;
      out[t] = src[k];
    }
  }
}

} // namespace fuse
} // namespace smesh

extern "C" int smesh_fuse_labels(const float* dist, int64_t P, int C, float dont_care_threshold, int32_t* labels_out,
                                 void* stream_v)
{
  if (P < 0 || C < 1 || (P > 0 && (!dist || !labels_out)))
  {
    set_error("smesh_fuse_labels: invalid argument");
    return SMESH_ERR_INVALID_ARGUMENT;
  }
  if (P > 0)
  {
    labels_kernel<<<(unsigned) ((P + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream_v)>>>(dist, P, C, dont_care_threshold,
                                                                                                labels_out);
    SMESH_LAUNCH_CHECK("labels_kernel");
  }
  return SMESH_OK;
}

extern "C" int smesh_fuse_render(const void* annotations, int64_t P, int elem_bytes, const uint32_t* ids32, int64_t n_pix,
                                 const void* background, void* out, void* stream_v)
{
  if (P < 0 || elem_bytes < 1 || n_pix < 0 || (n_pix > 0 && (!ids32 || !background || !out)) || (P > 0 && !annotations))
  {
    set_error("smesh_fuse_render: invalid argument");
    return SMESH_ERR_INVALID_ARGUMENT;
  }
  if (n_pix == 0)
  {
    return SMESH_OK;
  }
  const bool word_aligned = elem_bytes % 4 == 0 && ((reinterpret_cast<uintptr_t>(annotations) | reinterpret_cast<uintptr_t>(background) |
                                                      reinterpret_cast<uintptr_t>(out)) & 3) == 0;
  const int words = word_aligned ? elem_bytes / 4 : 0;
  const int64_t total = n_pix * (words > 0 ? words : elem_bytes);
  int64_t blocks = (total + 255) / 256;
  const int64_t cap = (int64_t) num_sms() * 16;
  if (blocks > cap) blocks = cap;
  gather_kernel<<<(unsigned) blocks, 256, 0, static_cast<cudaStream_t>(stream_v)>>>(
    static_cast<const unsigned char*>(annotations), P, words, elem_bytes, ids32, n_pix, static_cast<const unsigned char*>(background),
    static_cast<unsigned char*>(out));
  SMESH_LAUNCH_CHECK("gather_kernel");
  return SMESH_OK;
}
