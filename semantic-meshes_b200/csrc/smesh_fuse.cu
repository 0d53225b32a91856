// Label fusion for sm_100a: MeshAggregator.add / get (include/semantic_meshes/fusion/Mesh.h:57-133 with the aggregator
// chains of python/semantic_meshes/src/Fusion.cu:46-92).
//
// The reference copies the view to the host, builds a serial std::map histogram and then does a mutex-guarded vector add
// per pixel under OpenMP. Here one view is three launches on the caller's stream:
//   1. count_kernel   - per-face pixel count of this view (Mesh.h:90-93), runs of equal ids inside a warp merged into
//                       one atomicAdd; also writes the flat-order uint32 copy of the ids when the input is strided or
//                       not 32-bit
//   2. scatter_kernel - THE hot kernel, HBM-bound: the (n_pix, C) probability image is streamed exactly once through a
//                       multi-stage shared-memory ring filled by TMA bulk copies (cp.async.bulk + mbarrier, evict-first
//                       L2 policy so the stream does not flush the accumulator rows / ids / counters out of L2); each
//                       consumer warp owns 32 consecutive pixels of a stage: lane = pixel for the gate (sequential class
//                       sum, Mesh.h:98) and weight (Mesh.h:100-103), then lanes regroup as (run of equal face id, 4-class
//                       chunk) to reduce the run in registers and issue ONE 128-bit red.global.add.v4.f32 per chunk into
//                       the 16-byte padded accumulator row
//   3. clear_kernel   - zero the touched counters again (cheaper than a P-sized memset per view)
// No tensor cores: this is an irregular gather/scatter, not a contraction.
#include "smesh_common.cuh"

#include <math_constants.h>

namespace smesh {
namespace fuse {

constexpr uint32_t INVALID_ID = 0xFFFFFFFFu;

// ---------------------------------------------------------------------------------------------------------------------
// PTX helpers (mbarrier + bulk async copy = the non-tensor TMA path, SASS: UBLKCP / SYNCS)
// ---------------------------------------------------------------------------------------------------------------------

__device__ __forceinline__ uint32_t smem_u32(const void* p)
{
  return (uint32_t) __cvta_generic_to_shared(p);
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar)
{
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
  asm volatile(
    "{\n"
    ".reg .pred p;\n"
    "SMESH_WAIT_%=:\n"
    "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
    "@p bra SMESH_DONE_%=;\n"
    "bra SMESH_WAIT_%=;\n"
    "SMESH_DONE_%=:\n"
    "}\n" ::"r"(smem_u32(bar)),
    "r"(parity)
    : "memory");
}

__device__ __forceinline__ uint64_t l2_evict_first_policy()
{
  uint64_t policy;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
  return policy;
}

// global -> shared bulk copy, completion signalled on an mbarrier; bytes % 16 == 0, both addresses 16-byte aligned
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar, uint64_t policy)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
                 smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
               : "memory");
}

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

// mul aggregator input: LogProb(pow(p, w)) as -log (Fusion.cu:83-87, tt/numeric/LogProb.h:66-71); "zero" (isinf of
// either sign) is the absorbing +inf (LogProb.h:106-118).
__device__ __forceinline__ float neg_log_pow(float p, float w)
{
  const float q = powf(p, w);
  float l = (q == 0.0f) ? CUDART_INF_F : -logf(q);
  if (isinf(l))
  {
    l = CUDART_INF_F;
  }
  return l;
}

// ---------------------------------------------------------------------------------------------------------------------
// 1. / 3. per-face pixel count of one view and its reset
// ---------------------------------------------------------------------------------------------------------------------

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

template <typename IdT>
__global__ void __launch_bounds__(256) count_kernel(const IdT* __restrict__ ids, int64_t stride_outer, int64_t stride_inner,
                                                    int64_t n_inner, int64_t npix, int64_t P, uint32_t* __restrict__ counts,
                                                    uint32_t* __restrict__ ids32, int flat)
{
  const int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  uint32_t id = INVALID_ID;
  if (i < npix)
  {
    int64_t off = i;
    if (!flat)
    {
      const int64_t o = i / n_inner, in = i - o * n_inner;
      off = o * stride_outer + in * stride_inner;
    }
    id = sanitize_id(ids[off], P);
    if (ids32 != nullptr)
    {
      ids32[i] = id;
    }
  }
  // one atomic per run of equal ids inside the warp
  const uint32_t prev = __shfl_up_sync(0xFFFFFFFFu, id, 1);
  const bool head = (lane == 0) || (prev != id);
  const uint32_t headmask = __ballot_sync(0xFFFFFFFFu, head);
  if (head && id != INVALID_ID)
  {
    const uint32_t above = headmask & ~((2u << lane) - 1u);
    const int next = above ? (__ffs(above) - 1) : 32;
    atomicAdd(counts + id, (uint32_t) (next - lane));
  }
}

__global__ void __launch_bounds__(256) clear_kernel(const uint32_t* __restrict__ ids32, int64_t npix, int64_t P,
                                                    uint32_t* __restrict__ counts)
{
  const int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (i < npix)
  {
    const uint32_t id = ids32[i];
    if ((int64_t) id < P)
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
  const uint32_t* counts;  // [P]
  float* acc;              // [P][Cpad]
  int64_t npix;
  int64_t P;
  int64_t ntiles;
  int C, Cpad;
  int stages;
  float iew;
};

// Shared memory: [stages][NW*32*C] floats | per consumer warp {w[32], aux[32], run_span[32], run_id[32]} | full[stages], empty[stages]
template <int KIND, int CT>
__global__ void __launch_bounds__(288) scatter_kernel(ScatterArgs a)
{
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int C = CT > 0 ? CT : a.C;
  const int Cpad = CT > 0 ? ((CT + 3) & ~3) : a.Cpad;
  const int NW = (int) (blockDim.x >> 5) - 1; // consumer warps; warp 0 produces
  const int tile_px = NW * 32;
  const size_t stage_floats = (size_t) tile_px * C;
  const int stages = a.stages;

  float* stage_base = reinterpret_cast<float*>(smem_raw);
  uint32_t* scratch = reinterpret_cast<uint32_t*>(stage_base + stage_floats * stages);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(scratch + (size_t) NW * 128);
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

  if (warp == 0)
  {
    // ===== producer: one lane streams the tiles of this CTA into the ring =====
    if (lane == 0)
    {
      const uint64_t policy = l2_evict_first_policy();
      int it = 0;
      for (int64_t tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x, it++)
      {
        const int s = it % stages;
        const int use = it / stages;
        if (use > 0)
        {
          mbar_wait(empty_bar + s, (uint32_t) ((use - 1) & 1));
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
      }
    }
    return;
  }

  // ===== consumers =====
  const int cw = warp - 1;
  float* wbuf = reinterpret_cast<float*>(scratch + (size_t) cw * 128);
  uint32_t* run_span = scratch + (size_t) cw * 128 + 64;
  uint32_t* run_id = scratch + (size_t) cw * 128 + 96;
  const int nchunks = Cpad >> 2;

  int it = 0;
  for (int64_t tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x, it++)
  {
    const int s = it % stages;
    const uint32_t parity = (uint32_t) ((it / stages) & 1);
    const int64_t i = tile * tile_px + (int64_t) cw * 32 + lane;

    // independent of the probability tile: issue before waiting on it
    uint32_t id = INVALID_ID;
    float wt = 1.0f;
    uint32_t n = 1;
    if (i < a.npix)
    {
      id = __ldg(a.ids + i);
      if (a.weights != nullptr)
      {
        wt = __ldg(a.weights + i);
      }
    }
    const bool valid = (int64_t) id < a.P;
    if (valid)
    {
      n = __ldg(a.counts + id);
    }

    mbar_wait(full_bar + s, parity);
    const float* tile_s = stage_base + stage_floats * s + (size_t) cw * 32 * C;
    const float* row = tile_s + (size_t) lane * C;

    // ---- phase A: lane = pixel. Gate (Mesh.h:95-98): sequential float sum of the class vector > 0.5 ----
    float sum = 0.0f;
    float best = 0.0f;
    int best_c = 0;
    if (valid)
    {
      if (KIND == SMESH_KIND_SUMMAX)
      {
        best = row[0];
      }
#pragma unroll 8
      for (int c = 0; c < C; c++)
      {
        const float p = row[c];
        sum = __fadd_rn(sum, p);
        if (KIND == SMESH_KIND_SUMMAX && p > best) // first maximum, strict > (tt/tensor/util/ArgComp.h:3-18)
        {
          best = p;
          best_c = c;
        }
      }
    }
    const bool ok = valid && (sum > 0.5f);
    const float w = pixel_weight(a.iew, n, wt);

    if constexpr (KIND == SMESH_KIND_SUMMAX)
    {
      // one class per pixel (Fusion.cu:51-56): a single scalar reduction, no run merging needed
      if (ok)
      {
        red_add_f32(a.acc + (size_t) id * Cpad + best_c, __fmul_rn(best, w));
      }
    }
    else
    {
      // ---- runs of equal face id among consecutive accepted pixels of this warp ----
      const uint32_t key = ok ? id : INVALID_ID;
      const uint32_t prev = __shfl_up_sync(0xFFFFFFFFu, key, 1);
      const bool head = ok && (lane == 0 || prev != key);
      const uint32_t headmask = __ballot_sync(0xFFFFFFFFu, head);
      const uint32_t okmask = __ballot_sync(0xFFFFFFFFu, ok);
      wbuf[lane] = w;
      if (head)
      {
        const int r = __popc(headmask & ((1u << lane) - 1u));
        const uint32_t brk = (headmask | ~okmask) & ~((2u << lane) - 1u);
        const uint32_t end = brk ? (uint32_t) (__ffs(brk) - 1) : 32u;
        run_span[r] = (uint32_t) lane | (end << 8);
        run_id[r] = id;
      }
      __syncwarp();

      // ---- phase B: lane = (run, 4-class chunk) ----
      const int nitems = __popc(headmask) * nchunks;
      for (int item = lane; item < nitems; item += 32)
      {
        const int r = item / nchunks;
        const int c0 = (item - r * nchunks) << 2;
        const uint32_t span = run_span[r];
        const int p_begin = (int) (span & 0xFF), p_end = (int) (span >> 8);
        const bool h1 = c0 + 1 < C, h2 = c0 + 2 < C, h3 = c0 + 3 < C;
        float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f, a3 = 0.0f;
        for (int p = p_begin; p < p_end; p++)
        {
          const float pw = wbuf[p];
          const float* q = tile_s + (size_t) p * C + c0;
          if (KIND == SMESH_KIND_SUM)
          {
            // weighted::sum (tt/aggregator/MiscOps.h:83-93): acc += probs * w
            a0 = __fadd_rn(a0, __fmul_rn(q[0], pw));
            if (h1) a1 = __fadd_rn(a1, __fmul_rn(q[1], pw));
            if (h2) a2 = __fadd_rn(a2, __fmul_rn(q[2], pw));
            if (h3) a3 = __fadd_rn(a3, __fmul_rn(q[3], pw));
          }
          else
          {
            a0 = __fadd_rn(a0, neg_log_pow(q[0], pw));
            if (h1) a1 = __fadd_rn(a1, neg_log_pow(q[1], pw));
            if (h2) a2 = __fadd_rn(a2, neg_log_pow(q[2], pw));
            if (h3) a3 = __fadd_rn(a3, neg_log_pow(q[3], pw));
          }
        }
        red_add_v4(a.acc + (size_t) run_id[r] * Cpad + c0, a0, a1, a2, a3);
      }
    }
    __syncwarp();
    if (lane == 0)
    {
      mbar_arrive(empty_bar + s);
    }
  }
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
  const float w = pixel_weight(a.iew, a.counts[id], a.weights ? a.weights[i] : 1.0f);
  float* dst = a.acc + (size_t) id * a.Cpad;
  if (KIND == SMESH_KIND_SUMMAX)
  {
    red_add_f32(dst + best_c, __fmul_rn(best, w));
    return;
  }
  for (int c = 0; c < a.C; c++)
  {
    red_add_f32(dst + c, KIND == SMESH_KIND_SUM ? __fmul_rn(row[c], w) : neg_log_pow(row[c], w));
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// get(): per-face class distribution
// ---------------------------------------------------------------------------------------------------------------------

template <int KIND>
__global__ void __launch_bounds__(256) get_kernel(const float* __restrict__ acc, int64_t P, int C, int Cpad,
                                                  float* __restrict__ out)
{
  const int64_t r = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= P)
  {
    return;
  }
  const float* row = acc + (size_t) r * Cpad;
  float* o = out + (size_t) r * C;
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
      o[c] = v; // parked, rescaled below
    }
    norm = __fadd_rn(norm, fabsf(v));
  }
  const float inv = __fdiv_rn(1.0f, norm);
  for (int c = 0; c < C; c++)
  {
    const float v = (KIND == SMESH_KIND_MUL) ? o[c] : row[c];
    const float x = __fmul_rn(v, inv);
    o[c] = (isnan(x) || isinf(x)) ? 0.0f : x; // Fusion.h:79-95
  }
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
  if (C <= 24) { cfg = {8, 4}; return true; }
  if (C <= 48) { cfg = {4, 4}; return true; }
  if (C <= 96) { cfg = {4, 3}; return true; }
  if (C <= 192) { cfg = {4, 2}; return true; }
  if (C <= 400) { cfg = {2, 2}; return true; }
  if (C <= 800) { cfg = {1, 2}; return true; }
  return false;
}

static size_t ring_smem_bytes(int C, const RingConfig& cfg)
{
  return (size_t) cfg.stages * cfg.consumer_warps * 32 * C * 4 + (size_t) cfg.consumer_warps * 128 * 4 + (size_t) cfg.stages * 16;
}

template <int KIND, int CT>
static int launch_scatter_ring(const ScatterArgs& args_in, const RingConfig& cfg, cudaStream_t stream)
{
  ScatterArgs args = args_in;
  const size_t smem = ring_smem_bytes(args.C, cfg);
  auto kernel = scatter_kernel<KIND, CT>;
  static thread_local size_t configured_smem = 0;
  static thread_local int blocks_per_sm = 0;
  static thread_local int configured_threads = 0;
  static thread_local int configured_device = -1;
  const int threads = (cfg.consumer_warps + 1) * 32;
  int device = 0;
  SMESH_CUDA_CHECK(cudaGetDevice(&device));
  if (configured_smem != smem || configured_threads != threads || configured_device != device)
  {
    SMESH_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
    SMESH_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, kernel, threads, smem));
    if (blocks_per_sm < 1)
    {
      set_error("scatter_kernel does not fit on an SM (C=%d, %zu bytes of shared memory)", args.C, smem);
      return SMESH_ERR_UNSUPPORTED;
    }
    configured_smem = smem;
    configured_threads = threads;
    configured_device = device;
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

template <int KIND>
static int launch_scatter(const ScatterArgs& args, cudaStream_t stream)
{
  RingConfig cfg;
  const bool aligned = (reinterpret_cast<uintptr_t>(args.probs) & 15) == 0;
  if (aligned && ring_config(args.C, cfg))
  {
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
static int launch_count(const void* ids, int64_t so, int64_t si, int64_t n_inner, int64_t npix, int64_t P, uint32_t* counts,
                        uint32_t* ids32, bool flat, cudaStream_t stream)
{
  const int64_t blocks = (npix + 255) / 256;
  count_kernel<IdT><<<(unsigned) blocks, 256, 0, stream>>>(static_cast<const IdT*>(ids), so, si, n_inner, npix, P, counts,
                                                           ids32, flat ? 1 : 0);
  SMESH_LAUNCH_CHECK("count_kernel");
  return SMESH_OK;
}

static int add_view(int kind, const void* ids, int id_dtype, int64_t ids_so, int64_t ids_si, const float* probs,
                    const float* weights, int64_t w_so, int64_t w_si, int64_t n_outer, int64_t n_inner, int C, int64_t P,
                    float iew, uint32_t* counts, uint32_t* ids32, float* acc, cudaStream_t stream)
{
  const int64_t npix = n_outer * n_inner;
  if (npix == 0 || P == 0)
  {
    return SMESH_OK;
  }
  const bool ids_flat = (ids_si == 1 || n_inner == 1) && (ids_so == n_inner || n_outer == 1);
  // 32-bit ids already in flat order are consumed in place (int32: negative values read as >= 2^31 and fail `id < P`)
  const bool zero_copy = ids_flat && (id_dtype == SMESH_ID_U32 || (id_dtype == SMESH_ID_I32 && P <= 0x7FFFFFFFll));
  uint32_t* ids32_out = zero_copy ? nullptr : ids32;
  int rc;
  switch (id_dtype)
  {
    case SMESH_ID_U32: rc = launch_count<uint32_t>(ids, ids_so, ids_si, n_inner, npix, P, counts, ids32_out, ids_flat, stream); break;
    case SMESH_ID_I32: rc = launch_count<int32_t>(ids, ids_so, ids_si, n_inner, npix, P, counts, ids32_out, ids_flat, stream); break;
    case SMESH_ID_U64: rc = launch_count<uint64_t>(ids, ids_so, ids_si, n_inner, npix, P, counts, ids32_out, ids_flat, stream); break;
    case SMESH_ID_I64: rc = launch_count<int64_t>(ids, ids_so, ids_si, n_inner, npix, P, counts, ids32_out, ids_flat, stream); break;
    default: set_error("smesh_fuse_add: unknown id dtype %d", id_dtype); return SMESH_ERR_INVALID_ARGUMENT;
  }
  if (rc != SMESH_OK)
  {
    return rc;
  }
  const uint32_t* ids_flat_ptr = zero_copy ? static_cast<const uint32_t*>(ids) : ids32;

  if (weights != nullptr)
  {
    const bool w_flat = (w_si == 1 || n_inner == 1) && (w_so == n_inner || n_outer == 1);
    if (!w_flat)
    {
      set_error("smesh_fuse_add: the weights image must be contiguous in the same pixel order as the probability image");
      return SMESH_ERR_UNSUPPORTED;
    }
  }

  ScatterArgs args;
  args.probs = probs;
  args.ids = ids_flat_ptr;
  args.weights = weights;
  args.counts = counts;
  args.acc = acc;
  args.npix = npix;
  args.P = P;
  args.ntiles = 0;
  args.C = C;
  args.Cpad = smesh_fuse_padded_classes(C);
  args.stages = 0;
  args.iew = iew;
  switch (kind)
  {
    case SMESH_KIND_SUM: rc = launch_scatter<SMESH_KIND_SUM>(args, stream); break;
    case SMESH_KIND_SUMMAX: rc = launch_scatter<SMESH_KIND_SUMMAX>(args, stream); break;
    case SMESH_KIND_MUL: rc = launch_scatter<SMESH_KIND_MUL>(args, stream); break;
    default: set_error("smesh_fuse_add: unknown aggregator kind %d", kind); return SMESH_ERR_INVALID_ARGUMENT;
  }
  if (rc != SMESH_OK)
  {
    return rc;
  }
  clear_kernel<<<(unsigned) ((npix + 255) / 256), 256, 0, stream>>>(ids_flat_ptr, npix, P, counts);
  SMESH_LAUNCH_CHECK("clear_kernel");
  return SMESH_OK;
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
                              uint32_t* ids32, float* acc, void* stream)
{
  const int rc = check_add_args("smesh_fuse_add", kind, ids, probs, n_outer, n_inner, C, P, counts, ids32, acc);
  if (rc != SMESH_OK)
  {
    return rc;
  }
  return add_view(kind, ids, id_dtype, ids_stride_outer, ids_stride_inner, probs, weights, w_stride_outer, w_stride_inner,
                  n_outer, n_inner, C, P, iew, counts, ids32, acc, static_cast<cudaStream_t>(stream));
}

extern "C" int smesh_fuse_count(const void* ids, int id_dtype, int64_t ids_stride_outer, int64_t ids_stride_inner,
                                int64_t n_outer, int64_t n_inner, int64_t P, uint32_t* counts, uint32_t* ids32_out,
                                void* stream_v)
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
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  const bool flat = (ids_stride_inner == 1 || n_inner == 1) && (ids_stride_outer == n_inner || n_outer == 1);
  switch (id_dtype)
  {
    case SMESH_ID_U32: return launch_count<uint32_t>(ids, ids_stride_outer, ids_stride_inner, n_inner, npix, P, counts, ids32_out, flat, stream);
    case SMESH_ID_I32: return launch_count<int32_t>(ids, ids_stride_outer, ids_stride_inner, n_inner, npix, P, counts, ids32_out, flat, stream);
    case SMESH_ID_U64: return launch_count<uint64_t>(ids, ids_stride_outer, ids_stride_inner, n_inner, npix, P, counts, ids32_out, flat, stream);
    case SMESH_ID_I64: return launch_count<int64_t>(ids, ids_stride_outer, ids_stride_inner, n_inner, npix, P, counts, ids32_out, flat, stream);
    default: set_error("smesh_fuse_count: unknown id dtype %d", id_dtype); return SMESH_ERR_INVALID_ARGUMENT;
  }
}

extern "C" int smesh_fuse_scatter(int kind, const uint32_t* ids32, const float* probs, const float* weights, int64_t n_pix,
                                  int C, int64_t P, float iew, const uint32_t* counts, float* acc, void* stream_v)
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
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  ScatterArgs args;
  args.probs = probs;
  args.ids = ids32;
  args.weights = weights;
  args.counts = counts;
  args.acc = acc;
  args.npix = n_pix;
  args.P = P;
  args.ntiles = 0;
  args.C = C;
  args.Cpad = smesh_fuse_padded_classes(C);
  args.stages = 0;
  args.iew = iew;
  switch (kind)
  {
    case SMESH_KIND_SUM: return launch_scatter<SMESH_KIND_SUM>(args, stream);
    case SMESH_KIND_SUMMAX: return launch_scatter<SMESH_KIND_SUMMAX>(args, stream);
    default: return launch_scatter<SMESH_KIND_MUL>(args, stream);
  }
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
                                    int64_t P, float iew, uint32_t* counts, uint32_t* ids32, float* acc, void* stream)
{
  int rc = check_add_args("smesh_fuse_add_batch", kind, ids, probs, n_outer, n_inner, C, P, counts, ids32, acc);
  if (rc != SMESH_OK)
  {
    return rc;
  }
  if (B < 0)
  {
    set_error("smesh_fuse_add_batch: negative batch size");
    return SMESH_ERR_INVALID_ARGUMENT;
  }
  const size_t id_size = (id_dtype == SMESH_ID_U64 || id_dtype == SMESH_ID_I64) ? 8 : 4;
  for (int64_t b = 0; b < B; b++)
  {
    const void* ids_b = static_cast<const char*>(ids) + (size_t) b * ids_stride_view * id_size;
    const float* probs_b = probs + (size_t) b * probs_stride_view;
    const float* weights_b = weights ? weights + (size_t) b * w_stride_view : nullptr;
    rc = add_view(kind, ids_b, id_dtype, ids_stride_outer, ids_stride_inner, probs_b, weights_b, w_stride_outer,
                  w_stride_inner, n_outer, n_inner, C, P, iew, counts, ids32, acc, static_cast<cudaStream_t>(stream));
    if (rc != SMESH_OK)
    {
      return rc;
    }
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
  const int Cpad = smesh_fuse_padded_classes(C);
  const unsigned blocks = (unsigned) ((P + 255) / 256);
  switch (kind)
  {
    case SMESH_KIND_SUM: get_kernel<SMESH_KIND_SUM><<<blocks, 256, 0, stream>>>(acc, P, C, Cpad, out); break;
    case SMESH_KIND_SUMMAX: get_kernel<SMESH_KIND_SUMMAX><<<blocks, 256, 0, stream>>>(acc, P, C, Cpad, out); break;
    default: get_kernel<SMESH_KIND_MUL><<<blocks, 256, 0, stream>>>(acc, P, C, Cpad, out); break;
  }
  SMESH_LAUNCH_CHECK("get_kernel");
  return SMESH_OK;
}
