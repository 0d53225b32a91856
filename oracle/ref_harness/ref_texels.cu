// TEST INFRASTRUCTURE ONLY (never linked into the product library).
//
// Thin C driver around the GENUINE reference texel renderer: semantic_meshes::render::TexturedTriangleRenderer
// (include/semantic_meshes/render/TexturedTriangleRenderer.h) with the reference's own PLY loader and kernel, all
// compiled from /root/reference where they lie. The harness only supplies what the unbuildable Boost.Python layer
// (python/semantic_meshes/include/{Ply,Renderer,Camera}.h) would: the pixel struct, the cameras from raw arrays and
// the split of the {z, index} image into two planes.
// Built by oracle/Makefile into oracle/_ref/libref_texels.so (nvcc, sm_100a).
#include <template_tensors/TemplateTensors.h>
#include <semantic_meshes/data/Ply.h>
#include <semantic_meshes/render/TexturedTriangleRenderer.h>

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
  std::shared_ptr<semantic_meshes::render::TexturedTriangleRenderer> renderer;
};

thread_local std::string last_error;

// cam: 9 floats R (row-major), 3 floats t; intr: fx, fy, cx, cy (float values widened to double by the caller exactly
// like python/semantic_meshes/include/Camera.h:54); res: W, H
semantic_meshes::Camera make_camera(const float* R, const float* t, const double* f, const double* c, int W, int H)
{
  tt::Matrix3f rotation;
  for (int r = 0; r < 3; r++) for (int k = 0; k < 3; k++) rotation(r, k) = R[r * 3 + k];
  tt::Vector3f translation(t[0], t[1], t[2]);
  semantic_meshes::Camera camera;
  camera.intr = tt::geometry::projection::PinholeFC<tt::Vector2d, tt::Vector2d>(tt::Vector2d(f[0], f[1]), tt::Vector2d(c[0], c[1]));
  camera.extr = tt::geometry::transform::Rigid<float, 3>(rotation, translation);
  camera.resolution = tt::Vector2s((size_t) W, (size_t) H);
  return camera;
}

} // namespace

extern "C" const char* ref_texels_last_error()
{
  return last_error.c_str();
}

// n cameras: R [n][9], t [n][3], f [n][2], c [n][2], res [n][2]
extern "C" void* ref_texels_create(const char* ply_path, int n, const float* R, const float* t, const double* f, const double* c,
                                   const int* res, float texels_per_pixel)
{
  try
  {
    Handle* h = new Handle();
    h->ply = std::make_shared<semantic_meshes::data::Ply>(boost::filesystem::path(ply_path));
    std::vector<semantic_meshes::Camera> cameras;
    for (int i = 0; i < n; i++)
    {
      cameras.push_back(make_camera(R + 9 * i, t + 3 * i, f + 2 * i, c + 2 * i, res[2 * i], res[2 * i + 1]));
    }
    h->renderer = std::make_shared<semantic_meshes::render::TexturedTriangleRenderer>(h->ply, cameras, texels_per_pixel);
    return h;
  }
  catch (const std::exception& e)
  {
    last_error = e.what();
    return nullptr;
  }
}

extern "C" uint64_t ref_texels_primitives(void* h)
{
  return static_cast<Handle*>(h)->renderer->getPrimitivesNum();
}

// the faces as the constructor reordered them (TexturedTriangleRenderer.h:133-150), int32 [F][3]
extern "C" void ref_texels_faces(void* hv, int32_t* faces_out)
{
  Handle* h = static_cast<Handle*>(hv);
  auto& faces = h->ply->getTinyplyFaces();
  for (size_t k = 0; k < faces.size(); k++)
  {
    for (int j = 0; j < 3; j++) faces_out[3 * k + j] = faces[k](j);
  }
}

extern "C" int ref_texels_render(void* hv, const float* R, const float* t, const double* f, const double* c, int W, int H,
                                 uint32_t* idx_out, float* depth_out)
{
  try
  {
    Handle* h = static_cast<Handle*>(hv);
    semantic_meshes::Camera camera = make_camera(R, t, f, c, W, H);
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

extern "C" void ref_texels_destroy(void* h)
{
  delete static_cast<Handle*>(h);
}
