// TEST INFRASTRUCTURE ONLY (never linked into the product library).
//
// Thin C driver around the GENUINE reference aggregator: it includes the reference's own headers where
// they lie under /root/reference (include/semantic_meshes/fusion/Mesh.h + template-tensors) and only
// supplies (i) the aggregator chain compositions that the reference builds inside its Boost.Python module
// (python/semantic_meshes/src/Fusion.cu:46-92 - unbuildable here because Boost.Python is absent) and
// (ii) a plain-C entry point. Built by oracle/Makefile into oracle/_ref/libref_fusion.so.
#include <template_tensors/TemplateTensors.h>
#include <semantic_meshes/fusion/Mesh.h>

#include <cstdint>
#include <cstring>
#include <mutex>
#include <cmath>

namespace {

// Output maps of the chains; behaviour of python/semantic_meshes/include/Fusion.h:79-104.
struct finite_or_zero
{
  template <typename T, typename D = typename std::decay<T>::type>
  D operator()(T&& v) const volatile
  {
    return (math::isnan(v) || math::isinf(v)) ? static_cast<D>(0) : static_cast<D>(std::forward<T>(v));
  }
};

struct finite_or_zero_rowwise
{
  template <typename T, size_t N = tt::rows_v<T>::value, typename E = tt::decay_elementtype_t<T>>
  tt::VectorXT<E, N> operator()(T&& v) const volatile
  {
    return tt::elwise(finite_or_zero(), std::forward<T>(v));
  }
};

struct logprob_over_max
{
  template <typename T, size_t N = tt::rows_v<T>::value>
  tt::VectorXT<float, N> operator()(T&& p) const volatile
  {
    return tt::VectorXT<float, N>(tt::static_cast_to<float>(p / tt::max_el(p)));
  }
};

template <size_t C>
using LockedVec = atomic::Variable<tt::VectorXT<float, C>, atomic::op::Lock<std::mutex>>;
template <size_t C>
using LockedLogVec = atomic::Variable<tt::VectorXT<numeric::LogProb<float>, C>, atomic::op::Lock<std::mutex>>;

// Fusion.cu:66-76
template <size_t C>
auto make_sum()
{
  return aggregator::map_output(finite_or_zero_rowwise(),
    aggregator::map_output(tt::functor::normalize<tt::functor::l1_norm>(),
      aggregator::map_output(atomic::functor::load(),
        aggregator::weighted::sum<LockedVec<C>>())));
}

// Fusion.cu:46-64
template <size_t C>
auto make_summax()
{
  return aggregator::map_output(finite_or_zero_rowwise(),
    aggregator::map_output(tt::functor::normalize<tt::functor::l1_norm>(),
      aggregator::map_output(atomic::functor::load(),
        aggregator::map_input(
          [](auto&& probs_in, float weight){
            tt::VectorXT<float, C> probs_out(0);
            size_t max_index = tt::argmax<1>(probs_in)();
            probs_out(max_index) = probs_in(max_index);
            return tt::VectorXT<float, C>(probs_out * weight);
          },
          aggregator::sum<LockedVec<C>>()))));
}

// Fusion.cu:78-92
template <size_t C>
auto make_mul()
{
  return aggregator::map_output(finite_or_zero_rowwise(),
    aggregator::map_output(tt::functor::normalize<tt::functor::l1_norm>(),
      aggregator::map_output(logprob_over_max(),
        aggregator::map_output(atomic::functor::load(),
          aggregator::map_input(tt::functor::pow(),
            aggregator::prod<LockedLogVec<C>>())))));
}

struct AggBase
{
  virtual ~AggBase() {}
  virtual void add(int W, int H, const uint32_t* ids, const float* probs, const float* weights) = 0;
  virtual void get(float* out) = 0;
  virtual void reset() = 0;
};

template <size_t C, typename TChain>
struct AggImpl : AggBase
{
  using Agg = semantic_meshes::ModelAggregator<TChain, mem::alloc::host_heap>;
  Agg agg;
  uint64_t P;

  AggImpl(TChain chain, uint64_t P, float iew)
    : agg(P, iew, chain)
    , P(P)
  {
  }

  void add(int W, int H, const uint32_t* ids, const float* probs, const float* weights) override
  {
    const size_t npix = (size_t) W * (size_t) H;
    tt::AllocTensorT<uint32_t, mem::alloc::host_heap, tt::RowMajor, 2> prim(W, H);
    tt::AllocTensorT<float, mem::alloc::host_heap, tt::RowMajor, 3> pr(W, H, C);
    std::memcpy(prim.data(), ids, npix * sizeof(uint32_t));
    std::memcpy(pr.data(), probs, npix * C * sizeof(float));
    if (weights != nullptr)
    {
      tt::AllocTensorT<float, mem::alloc::host_heap, tt::RowMajor, 2> wt(W, H);
      std::memcpy(wt.data(), weights, npix * sizeof(float));
      agg.add(std::move(prim), tt::partial<2>(std::move(pr)), std::move(wt));
    }
    else
    {
      agg.add(std::move(prim), tt::partial<2>(std::move(pr)));
    }
  }

  void get(float* out) override
  {
    auto result = tt::eval<tt::RowMajor, mem::alloc::host_heap>(tt::total<1>(agg.get()));
    std::memcpy(out, result.data(), (size_t) P * C * sizeof(float));
  }

  void reset() override
  {
    agg.reset();
  }
};

template <size_t C, typename TChain>
AggBase* make(TChain chain, uint64_t P, float iew)
{
  return new AggImpl<C, TChain>(chain, P, iew);
}

template <size_t C>
AggBase* make_kind(int kind, uint64_t P, float iew)
{
  switch (kind)
  {
    case 0: return make<C>(make_sum<C>(), P, iew);
    case 1: return make<C>(make_summax<C>(), P, iew);
    case 2: return make<C>(make_mul<C>(), P, iew);
    default: return nullptr;
  }
}

} // namespace

// kind: 0 = sum, 1 = summax, 2 = mul. Returns NULL if the class count is not instantiated or kind unknown.
extern "C" void* ref_fusion_create(int kind, int C, uint64_t P, float iew)
{
  switch (C)
  {
    case 3: return make_kind<3>(kind, P, iew);
    case 19: return make_kind<19>(kind, P, iew);
    case 40: return make_kind<40>(kind, P, iew);
    default: return nullptr;
  }
}

// ids: uint32 [W][H]; probs: float [W][H][C]; weights: NULL or float [W][H] (layout of Fusion.h:42-64).
extern "C" void ref_fusion_add(void* h, int W, int H, const uint32_t* ids, const float* probs, const float* weights)
{
  static_cast<AggBase*>(h)->add(W, H, ids, probs, weights);
}

// out: float [P][C] = what MeshAggregator.get() returns (Fusion.h:72-76).
extern "C" void ref_fusion_get(void* h, float* out)
{
  static_cast<AggBase*>(h)->get(out);
}

extern "C" void ref_fusion_reset(void* h)
{
  static_cast<AggBase*>(h)->reset();
}

extern "C" void ref_fusion_destroy(void* h)
{
  delete static_cast<AggBase*>(h);
}

extern "C" int ref_fusion_has_classes(int C)
{
  return C == 3 || C == 19 || C == 40;
}
