// TEST INFRASTRUCTURE ONLY (never linked into the product library).
//
// Thin C driver around the GENUINE reference rasterizer: semantic_meshes::render::TriangleRenderer
// (include/semantic_meshes/render/TriangleRenderer.h) with the reference's own PLY loader
// (src/data/Ply.cpp, vendored tinyply) and kernel (template-tensors DeviceMutexRasterizer.h / Triangle.h),
// all compiled from /root/reference where they lie. The harness only supplies what the unbuildable
// Boost.Python layer (python/semantic_meshes/include/{Renderer,Camera}.h) would: the pixel struct, the camera
// construction from raw arrays, and the split of the {z, index} image into two planes.
// Built by oracle/Makefile into oracle/_ref/libref_raster.so (nvcc, sm_100a).
#include <template_tensors/TemplateTensors.h>
#include <semantic_meshes/data/Ply.h>
#include <semantic_meshes/render/TriangleRenderer.h>

#include <cstdint>
#include <memory>
#include <vector>
#include <string>

namespace {

// Same members as Renderer<T>::Pixel, python/semantic_meshes/include/Renderer.h:19-23
struct Pixel
{
  float z;
  uint32_t primitive_index;
};

struct Handle
{
  std::shared_ptr<semantic_meshes::data::Ply> ply;
  std::shared_ptr<semantic_meshes::render::TriangleRenderer> renderer;
};

thread_local std::string last_error;

} // namespace

extern "C" const char* ref_raster_last_error()
{
  return last_error.c_str();
}

extern "C" void* ref_raster_create(const char* ply_path)
{
  try
  {
    Handle* h = new Handle();
    h->ply = std::make_shared<semantic_meshes::data::Ply>(boost::filesystem::path(ply_path));
    h->renderer = std::make_shared<semantic_meshes::render::TriangleRenderer>(h->ply);
    return h;
  }
  catch (const std::exception& e)
  {
    last_error = e.what();
    return nullptr;
  }
}

extern "C" uint64_t ref_raster_primitives(void* h)
{
  return static_cast<Handle*>(h)->renderer->getPrimitivesNum();
}

// R row-major 3x3, t[3]: float (already rounded as Camera.h(py):19-52 does); f[2], c[2]: float values widened to
// double by the caller exactly like Camera.h(py):54. idx_out/depth_out: HOST arrays [W][H] (pixel (x,y) at x*H+y).
extern "C" int ref_raster_render(void* hv, const float* R, const float* t, const double* f, const double* c,
                                 int W, int H, uint32_t* idx_out, float* depth_out)
{
  try
  {
    Handle* h = static_cast<Handle*>(hv);
    tt::Matrix3f rotation;
    for (int r = 0; r < 3; r++) for (int k = 0; k < 3; k++) rotation(r, k) = R[r * 3 + k];
    tt::Vector3f translation(t[0], t[1], t[2]);
    tt::Vector2d focal(f[0], f[1]);
    tt::Vector2d principal(c[0], c[1]);

    semantic_meshes::Camera camera;
    camera.intr = tt::geometry::projection::PinholeFC<tt::Vector2d, tt::Vector2d>(focal, principal);
    camera.extr = tt::geometry::transform::Rigid<float, 3>(rotation, translation);
    camera.resolution = tt::Vector2s((size_t) W, (size_t) H);

    tt::AllocMatrixT<Pixel, mem::alloc::device, tt::RowMajor> image_d(camera.resolution);
    h->renderer->render(image_d, camera);

    std::vector<Pixel> host((size_t) W * H);
    cudaError_t err = cudaMemcpy(host.data(), image_d.data(), host.size() * sizeof(Pixel), cudaMemcpyDeviceToHost);
    if (err != cudaSuccess)
    {
      last_error = cudaGetErrorString(err);
      return 1;
    }
    for (size_t i = 0; i < host.size(); i++)
    {
      idx_out[i] = host[i].primitive_index;
      depth_out[i] = host[i].z;
    }
    return 0;
  }
  catch (const std::exception& e)
  {
    last_error = e.what();
    return 2;
  }
}

extern "C" void ref_raster_destroy(void* h)
{
  delete static_cast<Handle*>(h);
}
