// smesh_pipeline_views: the render + fuse loop of a whole batch of views enqueued by ONE call (include/smesh.h).
//
// What the reference's scripts do per view - `idx, depth = renderer.render(cam); aggregator.add(idx, probs)`
// (python/scripts/colorize_mesh.py:60-72) - costs a Python caller ~90 us of host time per view through the per-view entry
// points (argument marshalling, stream / event objects, allocations), more than the 78 us the GPU needs. Here the host
// side of the loop is native: two streams (the caller's for the fusion, a side stream owned by the library for the
// renders, up to `ring` views ahead), a ring of index images, two events per ring slot; ~10 driver calls per view.
#include "smesh_common.cuh"

#include <stdlib.h>

namespace smesh {

namespace {

struct RenderLane
{
  cudaStream_t side = nullptr;
  cudaEvent_t fork = nullptr, join = nullptr;
  cudaEvent_t rendered[SMESH_PIPELINE_MAX_RING] = {};
  cudaEvent_t added[SMESH_PIPELINE_MAX_RING] = {};
};

// One set per host thread and device (like the side stream of smesh_fuse_add_batch). NULL while `stream` is being
// captured and the set does not exist yet - nothing is created during a capture - or if creation fails.
RenderLane* render_lane(cudaStream_t stream)
{
  constexpr int MAX_DEVICES = 64;
  thread_local RenderLane lanes[MAX_DEVICES];
  int dev = -1;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= MAX_DEVICES)
  {
    cudaGetLastError();
    return nullptr;
  }
  RenderLane& l = lanes[dev];
  if (l.side == nullptr)
  {
    cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(stream, &st) != cudaSuccess || st != cudaStreamCaptureStatusNone)
    {
      cudaGetLastError();
      return nullptr;
    }
    RenderLane fresh;
    bool ok = cudaStreamCreateWithFlags(&fresh.side, cudaStreamNonBlocking) == cudaSuccess &&
              cudaEventCreateWithFlags(&fresh.fork, cudaEventDisableTiming) == cudaSuccess &&
              cudaEventCreateWithFlags(&fresh.join, cudaEventDisableTiming) == cudaSuccess;
    for (int k = 0; k < SMESH_PIPELINE_MAX_RING && ok; k++)
    {
      ok = cudaEventCreateWithFlags(&fresh.rendered[k], cudaEventDisableTiming) == cudaSuccess &&
           cudaEventCreateWithFlags(&fresh.added[k], cudaEventDisableTiming) == cudaSuccess;
    }
    if (!ok)
    {
      cudaGetLastError();
      for (int k = 0; k < SMESH_PIPELINE_MAX_RING; k++)
      {
        if (fresh.rendered[k]) cudaEventDestroy(fresh.rendered[k]);
        if (fresh.added[k]) cudaEventDestroy(fresh.added[k]);
      }
      if (fresh.fork) cudaEventDestroy(fresh.fork);
      if (fresh.join) cudaEventDestroy(fresh.join);
      if (fresh.side) cudaStreamDestroy(fresh.side);
      return nullptr;
    }
    l = fresh;
  }
  return &l;
}

} // namespace

} // namespace smesh

using namespace smesh;

extern "C" int smesh_pipeline_views(const void* mesh, size_t mesh_bytes, int64_t V, int64_t F, int64_t B, const float* R_host,
                                    const float* t_host, const double* f_host, const double* c_host, int W, int H,
                                    void* workspace, size_t workspace_bytes, int ring, uint32_t* idx_ring, float* depth_ring,
                                    int kind,
                                    const float* const* probs, const float* const* weights, int C, int64_t P, float iew,
                                    uint32_t* counts2, uint32_t count_epoch0, float* acc, void* stream_v)
{
  if (B < 0 || W < 1 || H < 1 || ring < 1 || ring > SMESH_PIPELINE_MAX_RING || (B > 0 && (!R_host || !t_host || !f_host || !c_host || !idx_ring || !probs || !counts2 || !acc)))
  {
    set_error("smesh_pipeline_views: invalid argument (B=%lld W=%d H=%d ring=%d or a null buffer)", (long long) B, W, H, ring);
    return SMESH_ERR_INVALID_ARGUMENT;
  }
  if (count_epoch0 == 0 || (int64_t) count_epoch0 + B - 1 > 255)
  {
    set_error("smesh_pipeline_views: count epochs %u .. %lld must stay inside 1 .. 255", count_epoch0,
              (long long) count_epoch0 + (long long) B - 1);
    return SMESH_ERR_INVALID_ARGUMENT;
  }
  if (B == 0)
  {
    return SMESH_OK;
  }
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  const size_t npix = (size_t) W * (size_t) H;
  RenderLane* lane = render_lane(stream);
  cudaStream_t rs = lane ? lane->side : stream; // without a side stream: the plain loop on the caller's stream
  if (lane)
  {
    SMESH_CUDA_CHECK(cudaEventRecord(lane->fork, stream));
    SMESH_CUDA_CHECK(cudaStreamWaitEvent(rs, lane->fork, 0));
  }
  int rc = SMESH_OK;
  for (int64_t b = 0; b < B && rc == SMESH_OK; b++)
  {
    const int k = (int) (b % ring);
    uint32_t* idx = idx_ring + (size_t) k * npix;
    float* depth = depth_ring ? depth_ring + (size_t) k * npix : nullptr;
    if (lane && b >= ring)
    {
      cudaStreamWaitEvent(rs, lane->added[k], 0); // view b - ring has been fused: its index image is free
    }
    rc = smesh_raster_render(mesh, mesh_bytes, V, F, R_host + 9 * b, t_host + 3 * b, f_host + 2 * b, c_host + 2 * b, W, H,
                             workspace, workspace_bytes, idx, depth, rs);
    if (rc != SMESH_OK)
    {
      break;
    }
    if (lane)
    {
      cudaEventRecord(lane->rendered[k], rs);
      cudaStreamWaitEvent(stream, lane->rendered[k], 0);
    }
    const uint32_t epoch = count_epoch0 + (uint32_t) b;
    // the index image is (W, H) with y fastest, like the probability image: outer = x, inner = y; 32-bit flat ids are
    // consumed in place, so the ids32 scratch is never touched (any non-null pointer satisfies the argument check)
    rc = smesh_fuse_add(kind, idx, SMESH_ID_U32, H, 1, probs[b], weights ? weights[b] : nullptr, H, 1, W, H, C, P, iew,
                        counts2 + (size_t) (epoch & 1u) * (size_t) P, epoch, idx, acc, stream);
    if (rc == SMESH_OK && lane)
    {
      cudaEventRecord(lane->added[k], stream);
    }
  }
  if (rc == SMESH_OK)
  {
    const cudaError_t err = cudaGetLastError(); // (the event calls above)
    if (err != cudaSuccess)
    {
      rc = cuda_fail(err, "smesh_pipeline_views (stream / event ordering)");
    }
  }
  if (lane)
  {
    // rejoin the side stream, also after an error: whatever was queued there is ordered before the caller's next work
    // (and a capture must not end with unjoined work)
    cudaEventRecord(lane->join, rs);
    cudaStreamWaitEvent(stream, lane->join, 0);
  }
  return rc;
}
