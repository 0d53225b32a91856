// TEST INFRASTRUCTURE ONLY (never linked into the product library).
//
// Thin C driver around the GENUINE reference PLY loader: semantic_meshes::data::Ply (src/data/Ply.cpp:9-15) with
// template-tensors' tinyply interface (tt/interface/tinyply/Tinyply.h:93-97,195-230) and the vendored tinyply, all compiled
// from /root/reference where they lie, on the CPU (g++). It pins semantic_meshes/data.py's loader: which files load, which
// are rejected, and the vertex / face arrays a loaded file yields. Built by oracle/Makefile into
// oracle/_ref/libref_ply.so.
#include <template_tensors/TemplateTensors.h>
#include <semantic_meshes/data/Ply.h>

#include <cstdint>
#include <cstring>
#include <memory>
#include <string>

namespace {
thread_local std::string last_error;
}

extern "C" const char* ref_ply_last_error()
{
  return last_error.c_str();
}

// -> handle, or NULL if the reference's loader throws (ref_ply_last_error() holds its message)
extern "C" void* ref_ply_load(const char* path)
{
  try
  {
    return new semantic_meshes::data::Ply(boost::filesystem::path(path));
  }
  catch (const std::exception& e)
  {
    last_error = e.what();
    return nullptr;
  }
}

extern "C" uint64_t ref_ply_vertices(void* h)
{
  return static_cast<semantic_meshes::data::Ply*>(h)->getTinyplyVertices().size();
}

extern "C" uint64_t ref_ply_faces(void* h)
{
  return static_cast<semantic_meshes::data::Ply*>(h)->getTinyplyFaces().size();
}

// verts_out float32[V][3], faces_out int32[F][3]
extern "C" void ref_ply_copy(void* h, float* verts_out, int32_t* faces_out)
{
  auto* ply = static_cast<semantic_meshes::data::Ply*>(h);
  auto& v = ply->getTinyplyVertices();
  auto& f = ply->getTinyplyFaces();
  for (size_t i = 0; i < v.size(); i++)
  {
    for (int k = 0; k < 3; k++) verts_out[3 * i + k] = v[i](k);
  }
  for (size_t i = 0; i < f.size(); i++)
  {
    for (int k = 0; k < 3; k++) faces_out[3 * i + k] = f[i](k);
  }
}

extern "C" void ref_ply_free(void* h)
{
  delete static_cast<semantic_meshes::data::Ply*>(h);
}
